from .swin import SwinTransformer, swin_t, swin_s, swin_b, swin_l

__all__ = ['SwinTransformer', 'swin_t', 'swin_s', 'swin_b', 'swin_l']
