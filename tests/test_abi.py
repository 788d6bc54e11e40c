"""The C-ABI library loads and exports every symbol include/dip.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import dip_engine as de
    if not os.path.exists(de.LIB_PATH):
        de.build()
    lib = ctypes.CDLL(de.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "dip.h")).read()
    declared = set(re.findall(r"\b(dip_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dip_plan_bind"} - declared  # no-op; keeps the set explicit
    assert declared == set(de.ABI_SYMBOLS), declared ^ set(de.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert de.lib().dip_version() == 100


def test_workspace_query_needs_no_gpu():
    import dip_engine as de
    desc = de.NetDesc(32, 3, 5, 128, 4, 1, 1, 0)
    n = de.lib().dip_plan_workspace_bytes(ctypes.byref(desc), 512, 512)
    assert 2 ** 30 < n < 12 * 2 ** 30
    bad = de.NetDesc(32, 3, 5, 60, 4, 1, 1, 0)   # widths must be multiples of 8
    assert de.lib().dip_plan_workspace_bytes(ctypes.byref(bad), 512, 512) == 0
    assert b"128" in de.lib().dip_last_error()


def test_workspace_query_rejects_unsupported_configurations_with_a_reason():
    """Host logic of the plan builder (no GPU): every rejected configuration returns 0 and sets dip_last_error()."""
    import dip_engine as de
    L = de.lib()

    def q(desc, H=512, W=512):
        return L.dip_plan_workspace_bytes(ctypes.byref(desc), H, W), L.dip_last_error().decode()

    ok = de.NetDesc(32, 3, 5, 128, 4, 1, 1, 0)
    base, _ = q(ok)
    assert base > 0
    for desc, H, W, word in [
        (de.NetDesc(32, 3, 5, 128, 8, 1, 1, 0), 512, 512, "num_channels_skip"),      # skip width other than 4 / 128
        (de.NetDesc(200, 3, 5, 128, 4, 1, 1, 0), 512, 512, "input depth"),           # > 128
        (de.NetDesc(32, 5, 5, 128, 4, 1, 1, 0), 512, 512, "num_output_channels"),
        (de.NetDesc(32, 3, 9, 128, 4, 1, 1, 0), 512, 512, "scales"),
        (de.NetDesc(32, 3, 5, 128, 4, 1, 1, 0), 500, 512, "divisible"),              # 500 % 32 != 0
    ]:
        n, err = q(desc, H, W)
        assert n == 0 and word in err, (n, err)
    # supported variants: inpainting (skip=128, nearest), other depths / sizes; bigger configurations need more memory
    wide, _ = q(de.NetDesc(32, 3, 5, 128, 128, 0, 1, 0))
    assert wide > base
    big, _ = q(ok, 1024, 1024)
    assert 3.5 * base < big < 4.5 * base
    small, _ = q(de.NetDesc(8, 1, 3, 128, 4, 0, 1, 1), 64, 96)
    assert 0 < small < base
    # flash-no-flash: image (3 channels) as input, per-scale upsampling modes; stored with 4 channels -> smaller than 32
    flash, _ = q(de.NetDesc(3, 3, 5, 128, 4, -1, 1, 0, 0b11100), 704, 768)
    assert 0 < flash
    # need_sigmoid = False is accepted; the input-gradient option adds the level-0 gradient buffers
    nosig, _ = q(de.NetDesc(32, 3, 5, 128, 4, 1, 0, 0))
    assert nosig == base
    ingrad, _ = q(de.NetDesc(32, 3, 5, 128, 4, 1, 1, 0, 0, 1))
    assert ingrad > base
    # per-scale widths (channels == 0: the arrays), precision bf16, downsample_mode 'avg'
    def per_scale(down, skip, in_ch=3, prec=0, dmode=0, up=None):
        d = de.NetDesc(in_ch, 3, len(down), 0, 0, 1, 1, prec)
        for i, (a, b, c) in enumerate(zip(down, up or down, skip)):
            d.channels_down[i], d.channels_up[i], d.channels_skip[i] = a, b, c
        d.downsample_mode = dmode
        return d
    snail, _ = q(per_scale([8, 16, 32, 64, 128], [0, 0, 0, 4, 4]))                       # denoising.ipynb c8:17-23
    assert 0 < snail < 0.25 * base
    kate, _ = q(per_scale([16, 32, 64, 128, 128], [0] * 5, in_ch=32, dmode=1))           # restoration.ipynb c7:28-36
    assert snail < kate < base
    bf16, _ = q(de.NetDesc(32, 3, 5, 128, 4, 1, 1, 2))                                   # + the bf16 twins
    assert base < bf16 < 1.35 * base
    for desc, word in [(per_scale([8, 12, 32, 64, 128], [0] * 5), "multiples of 8"),
                       (per_scale([8, 16, 32, 64, 128], [0, 0, 0, 128, 128]), "num_channels_skip"),
                       (per_scale([8, 16, 32, 64, 256], [0] * 5), "multiples of 8"),
                       (per_scale([16, 32, 64, 128, 128], [0] * 5, dmode=2), "downsample_mode"),
                       (de.NetDesc(32, 3, 5, 128, 4, 1, 1, 3), "precision")]:
        n, err = q(desc)
        assert n == 0 and word in err, (n, err)


def test_downsampler_output_size_matches_torch_conv_arithmetic():
    import dip_engine as de
    L = de.lib()
    for n, K, f, pad in [(384, 16, 4, 6), (576, 16, 4, 6), (1024, 16, 4, 6), (33, 7, 2, 3), (45, 16, 4, 0), (64, 32, 8, 12),
                         (15, 16, 4, 0)]:
        want = (n + 2 * pad - K) // f + 1 if n + 2 * pad >= K else 0
        assert L.dip_lanczos_down_out_size(n, K, f, pad) == want == de.down_out_size(n, K, f, pad)


def test_integration_md_stub_matches_the_header_struct():
    """INTEGRATION.md section 2 shows the ctypes binding a reference maintainer adds: its NetDesc must have exactly the
    fields of dip_net_desc in include/dip.h, in order (a shorter struct makes the library read past it)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "dip.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} dip_net_desc;", hdr, re.S).group(1)
    header_fields = re.findall(r"^\s*int\s+(\w+)(?:\[8\])?;", body, re.M)
    header_arrays = re.findall(r"^\s*int\s+(\w+)\[8\];", body, re.M)
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    stub = re.search(r"class NetDesc\(ctypes.Structure\):.*?_fields_ = \[(.*?)\]\n", doc, re.S).group(1)
    stub_fields = re.findall(r'"(\w+)"', stub)
    import dip_engine as de
    assert header_fields == stub_fields == [n for n, _ in de.NetDesc._fields_] and len(header_fields) == 14
    assert header_arrays == header_fields[10:13] == [n for n, t in de.NetDesc._fields_ if t is not ctypes.c_int]
    ctor = re.search(r"desc = NetDesc\((.*?)\)\s+#", doc).group(1)
    assert len([x for x in ctor.split(",") if x.strip()]) == 10   # the per-scale arrays and downsample_mode stay zero
