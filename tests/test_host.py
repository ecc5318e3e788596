"""CPU tests of the host logic and of the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(_built_library):
    hdr = open(os.path.join(ROOT, "include", "ldmseg_b200.h")).read()
    declared = set(re.findall(r"\b(ldmseg_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("ldmseg_igemm_params")
    assert len(declared) >= 25
    lib = ctypes.CDLL(_built_library)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ldmseg_b200.h but not exported"
    from ldmseg import _native as nat
    assert set(nat.EXPORTS) == declared
    abi = int(re.search(r"#define LDMSEG_ABI_VERSION (\d+)", hdr).group(1))
    assert abi >= 2 and nat.load().ldmseg_version() == abi


def test_igemm_params_struct_layout_matches_header():
    """ctypes mirror of ldmseg_igemm_params: field order / count as in the header."""
    from ldmseg import _native as nat
    hdr = open(os.path.join(ROOT, "include", "ldmseg_b200.h")).read()
    body = hdr[hdr.index("typedef struct ldmseg_igemm_params {"):hdr.index("} ldmseg_igemm_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"([a-z_0-9]+)\s*(?:\[[A-Z_]+\])?\s*[,;]", body)
    fields = [f[0] for f in nat.IgemmParams._fields_]
    assert fields == names and len(fields) == 47


def test_product_scheduler_host_logic_matches_reference(golden_dir):
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle.make_golden import SCHED_KW
    g = np.load(f"{golden_dir}/scheduler.npz")
    s = DDIMNoiseScheduler(**SCHED_KW)
    np.testing.assert_array_equal(s.alphas_cumprod.numpy(), g["alphas_cumprod"])
    np.testing.assert_array_equal(s.final_alpha_cumprod.numpy(), g["final_alpha_cumprod"])
    for n in (10, 50, 100):
        s.set_timesteps_inference(n)
        np.testing.assert_array_equal(s.timesteps.numpy(), g[f"timesteps_{n}"])
        assert s.steps_offset == 1000 // n - 1          # Q3: the config value is overwritten
    for name in ("linear", "squaredcos_cap_v2", "sigmoid"):
        np.testing.assert_array_equal(DDIMNoiseScheduler(**dict(SCHED_KW, beta_schedule=name)).alphas_cumprod.numpy(),
                                      g[f"alphas_cumprod_{name}"])
    eps, x = torch.from_numpy(g["eps"]), torch.from_numpy(g["x"])
    t = torch.tensor([999, 19])
    noisy = s.add_noise(x, eps.clone(), t)
    np.testing.assert_array_equal(noisy.numpy(), g["add_noise"])
    np.testing.assert_array_equal(s.remove_noise(noisy, eps, t).numpy(), g["remove_noise"])
    assert len(s) == 1000 and s.init_noise_sigma == 1.0 and "DDIMScheduler(" in str(s)
    with pytest.raises(NotImplementedError):
        DDIMNoiseScheduler(beta_schedule="nope")
    with pytest.raises(RuntimeError):                    # the hot call has no CPU fallback
        s.step(eps, s.timesteps[0], x)


def test_product_state_dict_keys_match_reference(golden_dir):
    from ldmseg.models import UNet, GeneralVAESeg
    from oracle.make_golden import TINY, VAE_KW
    g = np.load(f"{golden_dir}/unet_glue_tiny.npz")
    unet = UNet(**TINY)
    unet.remove_cross_attention()
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4)
    assert sorted(unet.state_dict().keys()) == list(g["state_keys"])
    assert sum(p.numel() for p in unet.parameters()) == int(g["n_params"])
    assert torch.all(unet.conv_in.weight[:, 4:] == 0) and unet.conv_in.weight.shape == (320, 12, 3, 3)
    gs = np.load(f"{golden_dir}/segvae.npz")
    vae = GeneralVAESeg(**VAE_KW)
    assert sorted(vae.state_dict().keys()) == list(gs["state_keys"])
    assert vae.downsample_factor == 8 and vae.interpolation_factor == 2
    with pytest.raises(RuntimeError):
        unet(torch.zeros(1, 12, 16, 16), torch.tensor(1), None)
    with pytest.raises(RuntimeError):
        vae.decode(torch.zeros(1, 4, 8, 8))
    with torch.device("meta"):
        full = UNet()
        full.remove_cross_attention()
        full.modify_encoder(in_channels=8, cond_channels=4)
    assert sum(p.numel() for p in full.parameters()) == 815_556_484


def test_output_dict_and_descriptors():
    from ldmseg.utils import OutputDict
    from ldmseg.models import get_image_descriptor_model, UNet
    from oracle.make_golden import TINY
    o = OutputDict(sample=1)
    o["x"] = 2
    assert o.sample == 1 and o.x == 2 and list(o.keys()) == ["sample", "x"]
    unet = UNet(**TINY)
    assert get_image_descriptor_model("remove", None, unet) == (None, None, None)
    assert all(tb.attn2 is None for b in unet.down_blocks if hasattr(b, "attentions")
               for a in b.attentions for tb in a.transformer_blocks)
    with pytest.raises(NotImplementedError):
        get_image_descriptor_model("dino_image", None, unet)


def test_weight_packing_is_a_valid_gemm_layout():
    """pack_* layouts reproduce F.conv2d / F.linear when multiplied against an explicit im2col (host logic)."""
    import torch.nn.functional as F
    from ldmseg import _pack as pk
    g = torch.Generator().manual_seed(0)
    w = torch.randn(24, 40, 3, 3, generator=g)
    x = torch.randn(2, 40, 8, 8, generator=g)
    ref = F.conv2d(x, w, padding=1).permute(0, 2, 3, 1).reshape(-1, 24)
    cols = F.unfold(x, 3, padding=1).reshape(2, 40, 9, 64).permute(0, 3, 2, 1)       # [n, pix, tap, c]
    cols = F.pad(cols, (0, 24)).reshape(2 * 64, 9 * 64)                              # channels padded to 64
    torch.testing.assert_close(cols @ pk.pack_conv3x3(w).t(), ref, rtol=1e-4, atol=1e-4)
    cols2 = F.unfold(x, 3, padding=1).reshape(2, 40, 9, 64).permute(0, 3, 2, 1).reshape(128, 360)
    torch.testing.assert_close(F.pad(cols2, (0, 24)) @ pk.pack_conv3x3_im2col(w).t(), ref, rtol=1e-4, atol=1e-4)
    wl, bl = torch.randn(64, 16, generator=g), torch.randn(64, generator=g)
    wi, bi = pk.interleave_geglu(wl, bl)
    y = torch.randn(5, 16, generator=g) @ wi.t() + bi
    h, gate = (torch.randn(5, 16, generator=torch.Generator().manual_seed(0)) * 0, None)
    r = y.reshape(5, 2, 2, 16)
    xin = torch.linalg.lstsq(wi, (y - bi).t()).solution.t()
    full = xin @ wl.t() + bl
    torch.testing.assert_close(r[:, :, 0].reshape(5, 32), full[:, :32], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(r[:, :, 1].reshape(5, 32), full[:, 32:], rtol=1e-3, atol=1e-3)
    wt, bt = pk.pack_convT2x2(torch.randn(8, 6, 2, 2, generator=g), torch.randn(6, generator=g))
    assert wt.shape == (24, 64) and bt.shape == (24,)
    # block-tiled layout consumed by the TMA weight map: [N/16][K/64][16][64], rows zero-padded to 16
    w2 = torch.randn(40, 192, generator=g)
    tp = pk.tile_pack(w2)
    assert tp.shape == (3, 3, 16, 64) and pk.TILE_ROWS == 16
    for n, k in ((0, 0), (17, 70), (39, 191), (25, 128)):
        assert tp[n // 16, k // 64, n % 16, k % 64] == w2[n, k]
    assert tp[2, :, 8:, :].abs().sum() == 0          # rows 40..47 are padding


def test_upsample2_phase_packing_equals_interpolate_plus_conv():
    """pack_upsample2_conv3x3 (ldmseg_igemm_params.upsample2): the four stacked 2x2 phase matrices, applied to the
    input pixels the kernel reads -- tap (a, b) of phase (py, px) at (y+py+a-1, x+px+b-1), zero outside --
    reproduce F.interpolate(nearest x2) + conv3x3 (diffusers Upsample2D) at the output pixels (2y+py, 2x+px)."""
    import torch.nn.functional as F
    from ldmseg import _pack as pk
    g = torch.Generator().manual_seed(1)
    n, c, h = 24, 40, 6
    w, x = torch.randn(n, c, 3, 3, generator=g), torch.randn(2, c, h, h, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    packed = pk.pack_upsample2_conv3x3(w)
    npad, cp = 32, 64
    assert packed.shape == (4 * npad, 4 * cp)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            wp = packed[(2 * py + px) * npad:(2 * py + px) * npad + n].reshape(n, 4, cp)
            assert wp[:, :, c:].abs().sum() == 0                       # channel padding
            acc = torch.zeros(2, n, h, h)
            for a in range(2):
                for b in range(2):
                    acc += torch.einsum("bchw,nc->bnhw", xp[:, :, py + a:py + a + h, px + b:px + b + h], wp[:, 2 * a + b, :c])
            out[:, :, py::2, px::2] = acc
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    assert packed[n:npad].abs().sum() == 0                             # row padding of phase 0
    # block-tiled as the TMA weight map reads it: phase p starts at 16-row block p * ceil(N / 16) (IgemmKParams::up_nblk)
    tp = pk.tile_pack(packed)
    assert tp.shape == (4 * npad // 16, 4 * cp // 64, 16, 64)
    for ph in range(4):
        assert torch.equal(tp[ph * (npad // 16), 1, 3, :c], packed[ph * npad + 3, cp:cp + c])   # row 3, tap 1


def test_tiling_heuristic():
    from ldmseg.engine.plan import choose_tiling
    bn, split, pair = choose_tiling(4096, 320, 45)
    assert bn in (64, 128, 160, 256) and split >= 1 and not pair
    bn, split, pair = choose_tiling(64, 1280, 360, allow_pair=True)   # 8x8 level at batch 1: weight streaming -> split-K
    assert split > 1 and ((64 + 127) // 128) * ((1280 + bn - 1) // bn) * split <= 148
    assert not pair                                   # a single 128-row tile cannot form a CTA pair
    bn, split, pair = choose_tiling(8 * 4096, 2560, 5)      # plenty of tiles: no split
    assert split == 1
    bn, split, pair = choose_tiling(8 * 4096, 320, 45, allow_pair=True)   # ingest-bound conv: pair halves the B staging
    assert pair and bn in (128, 160, 256) and split == 1
    # 320-wide pair tiles (two N = 160 MMAs per k-step): the multi-wave N = 320 / 640 convolutions of batch 8, measured
    # in profiles/r02_bench_bn320.log; never with split-K, never unless the caller allows them (GEGLU launches do not)
    from ldmseg.engine.plan import choose_tiling_ex
    for (m, n, kb) in ((8 * 4096, 320, 45), (8 * 4096, 320, 135), (8 * 1024, 640, 90), (8 * 1024, 640, 270),
                       (4 * 4096, 320, 90)):
        bn, split, pair, tail, clus = choose_tiling_ex(m, n, kb, allow_pair=True, allow_tail=True, allow_320=True)
        assert (bn, split, pair, clus) == (320, 1, True, False), (m, n, kb, bn, split, pair, tail)
        assert choose_tiling_ex(m, n, kb, allow_pair=True, allow_tail=True)[0] != 320
    for (m, n, kb) in ((8 * 256, 1280, 180), (2 * 4096, 320, 45), (4096, 320, 45), (64, 1280, 180)):
        bn, split, pair, tail, clus = choose_tiling_ex(m, n, kb, allow_pair=True, allow_tail=True, allow_320=True)
        assert bn != 320 or (pair and split == 1), (m, n, kb, bn, split, pair)
    assert choose_tiling_ex(8 * 256, 1280, 180, allow_pair=True, allow_tail=True, allow_320=True)[0] == 256
    # short-K projections (10 / 20 k-blocks) are epilogue-bound and lose with the wide tile: measured, kept narrow
    for (m, n, kb) in ((8 * 1024, 640, 10), (8 * 4096, 320, 20), (8 * 1024, 1920, 10)):
        assert choose_tiling_ex(m, n, kb, allow_pair=True, allow_tail=True, allow_320=True)[0] != 320


def test_tuned_tiling_table_is_legal():
    """ldmseg/engine/tuned_b200.json (tools/tune_tiling.py): every measured override is a launch the kernel accepts
    and the planner returns it; LDMSEG_TUNED=0 / use_tuned=False falls back to the cycle model."""
    import json
    from ldmseg.engine import plan
    path = os.path.join(ROOT, "latent-diffusion-segmentation_b200", "ldmseg", "engine", "tuned_b200.json")
    entries = json.load(open(path))["entries"]
    assert entries
    for key, (bn, split, pair, us_tuned, us_model) in entries.items():
        m, n, kb = (int(v) for v in key.split(","))
        m_tiles = (m + 127) // 128
        tiles = m_tiles * ((n + bn - 1) // bn)
        assert bn in (64, 128, 160, 256) and 1 <= split <= 16 and split <= kb
        assert split == 1 or tiles * split <= 148                 # split-K CTAs must be co-resident
        assert tiles * split * 128 * bn <= 16 * 1024 * 1024       # fits the plan's split-K workspace
        if pair:
            assert bn != 64 and m_tiles >= 2
        assert us_tuned < us_model
        assert plan.choose_tiling(m, n, kb, allow_pair=True) == (bn, split, bool(pair))
        assert plan.choose_tiling(m, n, kb, allow_pair=True, use_tuned=False) != (bn, split, bool(pair))


def test_layernorm_fold_algebra():
    """pk.fold_layernorm: LN(x) W^T + b == rstd * (x W'^T - mean * colsum) + c (host logic of the folded launch)."""
    import torch.nn.functional as F
    from ldmseg import _pack as pk
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(37, 320, generator=g) * 1.7 + 0.4).double()
    w, b = torch.randn(96, 320, generator=g).double(), torch.randn(96, generator=g).double()
    gam, bet = torch.randn(320, generator=g).double() * 0.3 + 1, torch.randn(320, generator=g).double() * 0.2
    ref = F.layer_norm(x, (320,), gam, bet, 1e-5) @ w.t() + b
    wf, c, colsum = pk.fold_layernorm(w.float(), b.float(), gam.float(), bet.float())
    mu = x.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    got = rstd * (x @ wf.double().t() - mu * wf.double().sum(1)[None]) + c.double()[None]
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
    # colsum is taken over the bf16-rounded weights: the mean term then cancels what the tensor cores accumulate
    torch.testing.assert_close(colsum, wf.to(torch.bfloat16).float().sum(1))
    # GEGLU interleave commutes with the fold
    wi, bi = pk.interleave_geglu(wf[:64], c[:64])
    assert wi.shape == (64, 320) and torch.equal(wi[:16], wf[:16]) and torch.equal(wi[16:32], wf[32:48])


def test_descriptor_variants_configure_the_unet():
    """descriptors.py:67-105: 'learnable' adds queries, 'clip_image' adds the 1024->768 projection; both keep
    cross-attention (the engine then needs encoder_hidden_states)."""
    from ldmseg.models import UNet
    from ldmseg.models import descriptors as D
    from oracle.make_golden import TINY
    unet = UNet(**TINY)
    assert D.get_image_descriptor_model("learnable", None, unet) == (None, None, None)
    assert tuple(unet.object_queries.weight.shape) == (128, 768)
    unet2 = UNet(**TINY)
    unet2.modify_encoder_hidden_state_proj(1024, 768)
    assert tuple(unet2.encoder_hid_proj.weight.shape) == (768, 1024)
    assert any(tb.attn2 is not None for b in unet2.down_blocks if hasattr(b, "attentions")
               for a in b.attentions for tb in a.transformer_blocks)
    # a tiny CLIP vision tower from config (no network): last_feat is [B, D, T] as the reference's wrapper returns
    m = D.make_vision_descriptor(False, name="", hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                                 num_attention_heads=2, image_size=224, patch_size=56, projection_dim=16)
    enc = D.image_descriptors(m.eval(), torch.rand(2, 3, 40, 40))
    assert enc.shape[0] == 4 and enc.shape[2] == 32
    mp = D.make_vision_descriptor(True, name="", hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                                  num_attention_heads=2, image_size=224, patch_size=56, projection_dim=16)
    assert tuple(D.image_descriptors(mp.eval(), torch.rand(1, 3, 40, 40)).shape) == (2, 1, 16)
