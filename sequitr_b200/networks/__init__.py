from .unet import UNet, UNet2D, UNet3D, ModeKeys, BRIDGE_TYPES, DEFAULT_FILTERS, DEFAULT_DROPOUT  # noqa: F401
