"""wlseg: B200-native (sm_100a) implementation of the segmentation hot path of
pmeletis/IV2019-boosting-semantic-segmentation-with-weak-labels behind the reference's own
surface (train.py / evaluate.py / predict.py flags, system_factory.SemanticSegmentation,
problem_definitions JSON).  Host code is Python on PyTorch (memory, streams, torch.distributed);
every device op is hand-written CUDA in libwlseg.so (include/wlseg.h).  No CPU fallback.
"""

__version__ = '0.1.0'
