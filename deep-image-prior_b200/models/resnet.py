"""Import-compatibility shim: `from models.resnet import ResNet` (inpainting.ipynb c3, restoration.ipynb c3 import it
next to `skip`, whether or not they build one).  The ResNet generator (reference: models/resnet.py:44-96) is outside the
accelerated hot path and is not provided: constructing it raises."""


class ResNet(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("dip-b200: the ResNet builder is outside the accelerated hot path (SURVEY.md section 8f) "
                                  "and is not provided; the skip network (models.skip / get_net(..., 'skip', ...)) is")
