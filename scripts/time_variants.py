"""it/s of the round-2 skip-net variants through the notebook-facing API (models.skip + optimize-style step) on the engine vs
the same module tree on stock torch.cuda + cuDNN (models.allow_torch_execution): snail (denoising.ipynb c8:13-23, 256x384) and
restoration kate (restoration.ipynb c7:28-36, 512x512)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-image-prior_b200"))
import torch
import models
import dip_engine as de

dtype = torch.cuda.FloatTensor
torch.backends.cudnn.benchmark = True
CASES = {
    "snail 256x384": dict(args=(3, 3), kw=dict(num_channels_down=[8, 16, 32, 64, 128], num_channels_up=[8, 16, 32, 64, 128],
                                               num_channels_skip=[0, 0, 0, 4, 4], upsample_mode="bilinear", need_sigmoid=True,
                                               need_bias=True, pad="reflection", act_fun="LeakyReLU"), hw=(256, 384)),
    "restoration-kate 512x512": dict(args=(32, 3), kw=dict(num_channels_down=[16, 32, 64, 128, 128], num_channels_up=[16, 32, 64, 128, 128],
                                                            num_channels_skip=[0, 0, 0, 0, 0], filter_size_down=3, filter_size_up=3,
                                                            filter_skip_size=1, upsample_mode="bilinear", downsample_mode="avg",
                                                            need_sigmoid=True, need_bias=True, pad="reflection"), hw=(512, 512)),
}
for name, c in CASES.items():
    H, W = c["hw"]
    res = {}
    for engine in (True, False):
        torch.manual_seed(0)
        net = models.skip(*c["args"], **c["kw"]).type(dtype)
        z0 = (torch.rand(1, c["args"][0], H, W) * 0.1).type(dtype)
        target = torch.rand(1, 3, H, W).type(dtype)
        mse = torch.nn.MSELoss().type(dtype)
        noise = z0.clone()
        if not engine:
            net._dip_spec, net._dip_why = None, "timing the stock torch modules"
            models.allow_torch_execution(True)
            opt = torch.optim.Adam(net.parameters(), lr=0.01)
        else:
            opt = de.FusedAdam(list(net.parameters()), lr=0.01)

        def step():
            opt.zero_grad()
            loss = mse(net(z0 + noise.normal_() * (1. / 30)), target)
            loss.backward()
            opt.step()
            return loss
        for _ in range(15):
            step()
        torch.cuda.synchronize()
        n = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        res["engine" if engine else "torch+cuDNN"] = n / (e0.elapsed_time(e1) / 1000.0)
        models.allow_torch_execution(False)
    print("%s: engine %.1f it/s | stock torch + cuDNN %.1f it/s | x%.2f" % (name, res["engine"], res["torch+cuDNN"],
                                                                          res["engine"] / res["torch+cuDNN"]), flush=True)
