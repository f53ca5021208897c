from .swin import SwinTransformer, swin_t, swin_s, swin_b, swin_l
from .resnet import ResNet, resnet50

__all__ = ['SwinTransformer', 'swin_t', 'swin_s', 'swin_b', 'swin_l', 'ResNet', 'resnet50']
