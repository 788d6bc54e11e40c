"""Import-compatibility shim: `from models.unet import UNet` (inpainting.ipynb c3, restoration.ipynb c3 import it next
to `skip`, whether or not they build one).  The UNet builder (reference: models/unet.py:32-192) is outside the
accelerated hot path and is not provided: constructing it raises."""


class UNet(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("dip-b200: the UNet builder is outside the accelerated hot path (SURVEY.md section 8f) "
                                  "and is not provided; the skip network (models.skip / get_net(..., 'skip', ...)) is")
