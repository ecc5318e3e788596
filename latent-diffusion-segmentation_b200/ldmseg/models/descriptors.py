"""`ldmseg.models.get_image_descriptor_model` -- drop-in for
/root/reference/ldmseg/models/descriptors.py:67-105.

Only the released configuration's branch is built: 'remove' strips cross-attention from the UNet
(descriptors.py:94-96).  The CLIP image/text conditioned variants need checkpoints that cannot be
fetched offline and the cross-attention kernels are a later scope row.
"""


def get_image_descriptor_model(descriptor_name, pretrained_model_path, unet):
    text_encoder = tokenizer = image_descriptor_model = None
    if descriptor_name == 'remove':
        unet.remove_cross_attention()
        print('Successfully removed cross attention layers from unet')
    elif descriptor_name in ('dino_image', 'mae'):
        raise NotImplementedError(f'{descriptor_name} is not yet supported')  # as in the reference (:78-88)
    elif descriptor_name in ('clip_image', 'clip_image_proj', 'learnable', 'none'):
        raise NotImplementedError(
            f"image_descriptors='{descriptor_name}' keeps cross-attention, which the B200 engine does not "
            "build yet; use the released configuration image_descriptors='remove'")
    else:
        raise AssertionError(descriptor_name)
    return image_descriptor_model, text_encoder, tokenizer
