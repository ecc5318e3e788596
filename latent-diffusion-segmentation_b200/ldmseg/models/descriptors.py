"""`ldmseg.models.get_image_descriptor_model` -- drop-in for
/root/reference/ldmseg/models/descriptors.py:67-105.

Chooses what the UNet's cross-attention attends to:

  'remove'          released configuration: cross-attention stripped from the UNet            (:94-96)
  'none'            CLIP text encoder + tokenizer (77 tokens x 768)                            (:98-103)
  'clip_image'      CLIP ViT-L/14 patch tokens (257 x 1024) + a 1024->768 projection in the UNet (:70-73)
  'clip_image_proj' CLIP ViT-L/14 projected image embedding (1 x 768)                           (:75-77)
  'learnable'       128 learned queries of width 768 inside the UNet                          (:90-92)
  'dino_image', 'mae'  raise NotImplementedError, as the reference does                        (:78-88)

The CLIP encoders are the `transformers` library models the reference itself uses; they run once per batch,
outside the denoising loop.  What IS on the hot path -- the cross-attention of the 16 transformer blocks over
these tokens, with the doubled batch and the guidance combine of `sample()` -- runs on the CUDA engine
(`ldmseg_cross_attention`, `ldmseg_sampler_step` with cfg = 1).  Checkpoints cannot be fetched here (no network):
a name that is not a local directory yields the published architecture with random-init weights, like
`UNet.from_pretrained`.
"""
import os
from typing import Optional

import torch


def _clip():
    try:
        import transformers  # noqa: F401
        from transformers import (CLIPTextConfig, CLIPTextModel, CLIPTokenizer, CLIPVisionConfig, CLIPVisionModel,
                                  CLIPVisionModelWithProjection)
    except Exception as e:  # pragma: no cover - transformers is part of the reference's environment
        raise ImportError("the CLIP-conditioned descriptor variants need the `transformers` package") from e
    return CLIPTextConfig, CLIPTextModel, CLIPTokenizer, CLIPVisionConfig, CLIPVisionModel, CLIPVisionModelWithProjection


# ViT-L/14 (openai/clip-vit-large-patch14) and the SD-v1 text encoder (ViT-L/14 text tower), as published
_VIT_L14 = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                image_size=224, patch_size=14, projection_dim=768)
_TEXT_L14 = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                 num_attention_heads=12, max_position_embeddings=77, projection_dim=768)


def _load_or_init(cls, config_cls, name: str, default_cfg: dict, subfolder: Optional[str] = None, **overrides):
    path = os.path.join(name, subfolder) if (subfolder and name) else name
    if path and os.path.isdir(path):
        return cls.from_pretrained(path)
    cfg = dict(default_cfg)
    cfg.update(overrides)
    print(f"[ldmseg_b200] '{name}' is not a local directory: building {cls.__name__} from its published config "
          "with random-init weights (no network access)")
    return cls(config_cls(**cfg))


def make_vision_descriptor(with_projection: bool = False, name: str = "openai/clip-vit-large-patch14", **cfg):
    """MyCLIPVisionModel / MyCLIPVisionModelWithProjection of the reference (descriptors.py:15-56): forward
    returns {'last_feat': [B, D, T]} -- the caller permutes it to [B, T, D] (trainers_ldm_cond.py:1104)."""
    _, _, _, CLIPVisionConfig, CLIPVisionModel, CLIPVisionModelWithProjection = _clip()

    class MyCLIPVisionModel(CLIPVisionModel):
        def forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None):
            out = self.vision_model(pixel_values=pixel_values, output_attentions=output_attentions,
                                    output_hidden_states=output_hidden_states, return_dict=True)
            return {'last_feat': out.last_hidden_state.permute(0, 2, 1)}

    class MyCLIPVisionModelWithProjection(CLIPVisionModelWithProjection):
        def forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None):
            out = self.vision_model(pixel_values=pixel_values, output_attentions=output_attentions,
                                    output_hidden_states=output_hidden_states, return_dict=True)
            return {'last_feat': self.visual_projection(out.pooler_output).unsqueeze(-1)}

    cls = MyCLIPVisionModelWithProjection if with_projection else MyCLIPVisionModel
    return _load_or_init(cls, CLIPVisionConfig, name, _VIT_L14, **cfg)


def get_image_descriptor_model(descriptor_name, pretrained_model_path, unet):
    text_encoder = tokenizer = image_descriptor_model = None
    if descriptor_name == 'clip_image':
        image_descriptor_model = make_vision_descriptor(False)
        unet.modify_encoder_hidden_state_proj(1024, 768)
    elif descriptor_name == 'clip_image_proj':
        image_descriptor_model = make_vision_descriptor(True)
    elif descriptor_name == 'dino_image':
        raise NotImplementedError('DINO is not yet supported')            # as in the reference (:78-82)
    elif descriptor_name == 'mae':
        raise NotImplementedError('MAE is not yet supported')             # as in the reference (:84-88)
    elif descriptor_name == 'learnable':
        # the reference calls a misspelled method here (descriptors.py:91 vs unet.py:38); the intent is
        unet.define_learnable_embedding(128, 768)
        print(f'Successfully added learnable object queries to unet as {unet.object_queries}')
    elif descriptor_name == 'remove':
        unet.remove_cross_attention()
        print('Successfully removed cross attention layers from unet')
    else:
        assert descriptor_name == 'none'
        CLIPTextConfig, CLIPTextModel, CLIPTokenizer, _, _, _ = _clip()
        tok_dir = os.path.join(pretrained_model_path or "", "tokenizer")
        if os.path.isdir(tok_dir):
            tokenizer = CLIPTokenizer.from_pretrained(tok_dir)
        else:
            print(f"[ldmseg_b200] no tokenizer under '{pretrained_model_path}': prompts must be tokenised by the "
                  "caller (no network access)")
        text_encoder = _load_or_init(CLIPTextModel, CLIPTextConfig, pretrained_model_path or "", _TEXT_L14,
                                     subfolder="text_encoder")
        print('Succesfully loaded pretrained CLIP text encoder')
    return image_descriptor_model, text_encoder, tokenizer


@torch.no_grad()
def image_descriptors(image_descriptor_model, rgb_images: torch.Tensor) -> torch.Tensor:
    """`norm_resize_images` + descriptor forward of `sample()` (trainers_ldm_cond.py:663-675, 1102-1106):
    RGB in [0,1] -> [2B, T, D] encoder_hidden_states (the same descriptors for both halves of the doubled batch)."""
    import torch.nn.functional as F
    x = F.interpolate(rgb_images.float(), size=(224, 224), mode='bilinear', align_corners=False)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=x.device).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=x.device).view(1, 3, 1, 1)
    d = image_descriptor_model((x - mean) / std)['last_feat']
    d = d.view(d.shape[0], d.shape[1], -1).permute(0, 2, 1)
    return torch.cat([d] * 2).to(torch.float)
