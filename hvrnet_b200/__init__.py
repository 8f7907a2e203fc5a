"""hvrnet_b200: B200-native (sm_100a) implementation of HVRNet's per-key-frame inference hot path
behind the reference's Registry / config plugin surface.  Importing the package registers the
modules (ResNet, ResLayer, RPNHead, SingleRoIExtractor, HRNMPBBoxHead, SelsaBBoxHead,
SharedFCBBoxHead, HNMBRCNN, SelsaRCNN, FasterRCNN) like ``import mmdet.models`` does."""
from . import models  # noqa: F401  (registers the plugin classes)
from .builder import build_detector  # noqa: F401
from .config import Config  # noqa: F401

__version__ = '0.1.0'
