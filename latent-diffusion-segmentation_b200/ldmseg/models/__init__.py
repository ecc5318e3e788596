from .unet import UNet
from .vae import GeneralVAESeg, GeneralVAEImage
from .descriptors import get_image_descriptor_model
from .upscaler import Upscaler

__all__ = ['UNet', 'GeneralVAESeg', 'GeneralVAEImage', 'get_image_descriptor_model', 'Upscaler']
