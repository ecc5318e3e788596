"""GPU parity tests (-m gpu) at the BENCHMARKED configurations and for the round-2 rows of SURVEY.md §8(f).

Every measured error is written to gpurun_out/r02_parity.json (copied to profiles/ after a GPU round), so the
numbers behind the tolerances are on record.  Tolerances: bf16 operands (eps = 2^-8) with fp32 accumulation against
an fp32 oracle -- relative L2 per forward ~8e-3 (stated bar 2e-2); it compounds over a 50-step loop, where the
bar is stated per test; integer outputs (ids, segment tables) are compared by agreement or bit-exactly.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARITY_OUT = os.environ.get("LDMSEG_PARITY_OUT", os.path.join(ROOT, "gpurun_out", "r02_parity.json"))


def record(name, **values):
    try:
        os.makedirs(os.path.dirname(PARITY_OUT), exist_ok=True)
        data = {}
        if os.path.exists(PARITY_OUT):
            with open(PARITY_OUT) as f:
                data = json.load(f)
        data[name] = {k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in values.items()}
        with open(PARITY_OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print(f"[parity] {name}: {values}")


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def unets(dev):
    """(oracle UNet, product UNet on the GPU): SD-v1 width, cross-attention removed, 12-channel conv_in, same weights."""
    from ldmseg.models import UNet
    from oracle import ldmseg_restated as orc
    oracle_unet = orc.build_ldmseg_unet(seed=0, cond_channels=4, image_init="zero")
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        oracle_unet.conv_in.weight[:, 4:].copy_(torch.randn(oracle_unet.conv_in.weight[:, 4:].shape, generator=g) * 0.05)
    unet = UNet()
    unet.remove_cross_attention()
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4)
    unet.load_state_dict(oracle_unet.state_dict(), strict=True)
    return oracle_unet, unet.to(dev)


# a narrow UNet whose head dims (40 / 80) the attention kernel supports: cheap oracle runs for the state tests
SMALL = dict(block_out_channels=(320, 320, 640, 640), attention_head_dim=8)


@pytest.fixture(scope="module")
def vaes(dev):
    """(oracle AutoencoderKL, oracle seg VAE, product GeneralVAEImage, product GeneralVAESeg), same weights."""
    from ldmseg.models import GeneralVAEImage, GeneralVAESeg
    from oracle import diffusers_restated as dr
    from oracle import ldmseg_restated as orc
    from oracle.make_golden import VAE_KW
    torch.manual_seed(1)
    ref_vi = dr.AutoencoderKL().eval()
    kw = dict(VAE_KW, scaling_factor=0.18215)
    ref_vs = orc.GeneralVAESeg(**{k: v for k, v in kw.items() if k != "parametrization"}).eval()
    vi = GeneralVAEImage()
    vi.load_state_dict(ref_vi.state_dict(), strict=True)
    vi.set_scaling_factor(0.18215)
    vs = GeneralVAESeg(**kw)
    vs.load_state_dict(ref_vs.state_dict(), strict=True)
    return ref_vi, ref_vs, vi.to(dev), vs.to(dev)


# ------------------------------------------------------------------------------------------------ configs[1]
def test_config1_full_chain_50_steps(dev, unets, vaes):
    """BASELINE configs[1] as quoted: 512x512 RGB -> AutoencoderKL encode -> 50-step DDIM at a 64x64 latent
    (self-conditioning, batch 1, seed 42) -> seg decode -> argmax ids; product path vs the oracle chain
    (oracle/ldmseg_restated.py encode_inputs / sample / decode_latents).  Bars: rgb latents rel-L2 <= 2e-2,
    final latents rel-L2 <= 1.5e-1 (the per-forward ~8e-3 compounds over 50 self-conditioned steps; measured value
    in r02_parity.json), argmax-id agreement >= 0.90."""
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle import ldmseg_restated as orc
    from oracle.make_golden import SCHED_KW
    oracle_unet, unet = unets
    ref_vi, ref_vs, vi, vs = vaes
    rgb = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        ref_rgb_lat = orc.encode_inputs(rgb, ref_vi, 0.18215)
        ref_lat = orc.sample(oracle_unet, orc.DDIMNoiseScheduler(**SCHED_KW), ref_rgb_lat, 50, seed=42)
        ref_logits = orc.decode_latents(ref_lat, ref_vs)
    ref_ids = ref_logits.argmax(1)
    sampler = B200Sampler(unet, DDIMNoiseScheduler(**SCHED_KW), vi, vs, self_condition=True)
    x = rgb.to(dev)
    rgb_lat = sampler.encode_rgb(x)
    lat = sampler.sample(rgb_lat, 50, seed=42)
    ids, prob = sampler.generate(x, 50, seed=42)
    e_rgb, e_lat = rel_l2(rgb_lat, ref_rgb_lat), rel_l2(lat, ref_lat)
    agree = (ids.cpu().long() == ref_ids).float().mean().item()
    # the sampler alone, fed the oracle's rgb latents (isolates the loop from the encoder's error)
    lat2 = sampler.sample(ref_rgb_lat.to(dev), 50, seed=42)
    e_lat2 = rel_l2(lat2, ref_lat)
    # decode alone, fed the oracle's final latents
    ids3, _ = vs._get_engine().decode_ids(ref_lat.to(dev), scale=1.0 / vs.scaling_factor)
    agree3 = (ids3.cpu().long() == ref_ids).float().mean().item()
    record("config1_chain_b1_64x64_50steps", rgb_latents_rel_l2=e_rgb, final_latents_rel_l2=e_lat,
           final_latents_rel_l2_given_oracle_rgb=e_lat2, argmax_id_agreement=agree,
           argmax_id_agreement_decode_only=agree3)
    assert ids.shape == (1, 512, 512)
    assert e_rgb <= 2e-2 and e_lat <= 1.5e-1 and agree >= 0.90 and agree3 >= 0.99


def test_unet_forward_batch8_and_latent128(dev, unets):
    """The forward at the batch the CTA-pair plan is chosen for (B=8 @ 64x64) and at B=2 @ 128x128
    (configs[4] per-GPU share): rel-L2 <= 2e-2 vs the fp32 oracle."""
    oracle_unet, unet = unets
    g = torch.Generator().manual_seed(31)
    out = {}
    for shape, t in (((8, 12, 64, 64), 499), ((2, 12, 128, 128), 259)):
        x = torch.randn(*shape, generator=g)
        tt = torch.tensor(t)
        y = unet(x.to(dev), tt.to(dev), encoder_hidden_states=None).sample
        with torch.no_grad():
            ref = oracle_unet(x, tt).sample
        r = rel_l2(y, ref)
        out[f"b{shape[0]}_{shape[2]}"] = r
        assert torch.isfinite(y).all() and r <= 2e-2, (shape, r)
    plan = unet._get_engine().plan(8, 64)
    n_pair = sum(":p1" in t for t in plan.tags)
    record("unet_forward", rel_l2_b8_64=out["b8_64"], rel_l2_b2_128=out["b2_128"], pair_launches_b8=n_pair)
    assert n_pair > 0        # the batch-8 plan really runs CTA-pair launches


def test_image_encoder_512(dev, vaes):
    """AutoencoderKL encoder at the benchmarked 512x512 (batch 2): posterior mean rel-L2 <= 2e-2."""
    ref_vi, _, vi, _ = vaes
    x = torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(2)) * 2 - 1
    with torch.no_grad():
        ref = ref_vi.encode(x).latent_dist.mode()
    out = vi.encode(x.to(dev)).latent_dist.mode()
    r = rel_l2(out, ref)
    record("image_encoder_512", rel_l2=r)
    assert out.shape == ref.shape and r <= 2e-2


def test_fp32_residual_stream_ab(dev, unets):
    """A/B of the fp32 residual / skip stream (SURVEY.md §7 hard part 3) against the default bf16 stream: one
    forward at [1,12,64,64] and the 50-step loop at a 32x32 latent, both vs the fp32 oracle.  Recorded, and both
    variants held to the same bars."""
    from ldmseg.engine import plan as plan_mod
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle import ldmseg_restated as orc
    from oracle.make_golden import SCHED_KW
    oracle_unet, unet = unets
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 12, 64, 64, generator=g)
    tt = torch.tensor(999)
    rgb = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(3)) * 0.7
    with torch.no_grad():
        ref = oracle_unet(x, tt).sample
        ref_loop = orc.sample(oracle_unet, orc.DDIMNoiseScheduler(**SCHED_KW), rgb, 50, seed=42)
    res = {}
    old = plan_mod.RESID_F32
    try:
        for mode in (False, True):
            plan_mod.RESID_F32 = mode
            unet.invalidate_engine()
            y = unet(x.to(dev), tt.to(dev), encoder_hidden_states=None).sample
            s = B200Sampler(unet, DDIMNoiseScheduler(**SCHED_KW), self_condition=True)
            lat = s.sample(rgb.to(dev), 50, seed=42)
            res[mode] = (rel_l2(y, ref), rel_l2(lat, ref_loop))
    finally:
        plan_mod.RESID_F32 = old
        unet.invalidate_engine()
    record("fp32_residual_stream_ab", forward_rel_l2_bf16_stream=res[False][0], forward_rel_l2_f32_stream=res[True][0],
           loop50_rel_l2_bf16_stream=res[False][1], loop50_rel_l2_f32_stream=res[True][1])
    for mode in (False, True):
        assert res[mode][0] <= 2e-2 and res[mode][1] <= 1.5e-1, res


# ------------------------------------------------------------------------------------------------ sampler state
def test_sampler_state_follows_weights_and_scheduler(dev):
    """The cached plan / time-embedding table / graph are rebuilt when the UNet's weights or the scheduler change
    (ADVICE round 1), and the scheduler's prediction_type / clip_sample reach the fused step."""
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.models import UNet
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle import ldmseg_restated as orc
    from oracle.make_golden import SCHED_KW
    torch.manual_seed(0)
    o1 = orc.UNet(**SMALL)
    o1.remove_cross_attention()
    o1.modify_encoder(in_channels=8, cond_channels=4, init_mode_image="copy")
    with torch.no_grad():
        o1.conv_in.weight[:, 8:].normal_(0, 0.05)
    torch.manual_seed(1)
    o2 = orc.UNet(**SMALL)
    o2.remove_cross_attention()
    o2.modify_encoder(in_channels=8, cond_channels=4, init_mode_image="copy")
    unet = UNet(**SMALL)
    unet.remove_cross_attention()
    unet.modify_encoder(in_channels=8, cond_channels=4)
    unet.load_state_dict(o1.state_dict(), strict=True)
    unet = unet.to(dev)
    rgb = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(5)) * 0.5
    sch = DDIMNoiseScheduler(**SCHED_KW)
    s = B200Sampler(unet, sch, self_condition=True)
    a = s.sample(rgb.to(dev), 8, seed=3)
    ra = orc.sample(o1.eval(), orc.DDIMNoiseScheduler(**SCHED_KW), rgb, 8, seed=3)
    assert rel_l2(a, ra) <= 5e-2
    unet.load_state_dict({k: v.to(dev) for k, v in o2.state_dict().items()}, strict=True)     # new weights, same sampler
    b = s.sample(rgb.to(dev), 8, seed=3)
    rb = orc.sample(o2.eval(), orc.DDIMNoiseScheduler(**SCHED_KW), rgb, 8, seed=3)
    e_b = rel_l2(b, rb)
    assert e_b <= 5e-2 and rel_l2(b, ra) > 5 * e_b
    # scheduler variants through the fused step
    errs = {}
    for kw in (dict(prediction_type="sample"), dict(prediction_type="v_prediction"), dict(clip_sample=True, clip_sample_range=1.0)):
        s.scheduler = DDIMNoiseScheduler(**dict(SCHED_KW, **kw))
        c = s.sample(rgb.to(dev), 8, seed=3)
        rc = orc.sample(o2.eval(), orc.DDIMNoiseScheduler(**dict(SCHED_KW, **kw)), rgb, 8, seed=3)
        errs[str(kw)] = rel_l2(c, rc)
        assert errs[str(kw)] <= 8e-2, (kw, errs)     # clipping saturates x0: sign flips at the clip boundary cost more
    record("sampler_state_and_scheduler_variants", reload_rel_l2=e_b, **{f"rel_l2 {k}": v for k, v in errs.items()})


def test_extensions_run_in_graph_and_seeded(dev, unets):
    """Inpainting / DDPM replay the captured graph (device-resident per-step tables) and the DDPM noise follows the
    caller's seed: different seeds give different samples, equal seeds equal samples."""
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.schedulers import DDIMNoiseScheduler
    from ldmseg import _native as nat
    from oracle.make_golden import SCHED_KW
    _, unet = unets
    g = torch.Generator().manual_seed(9)
    rgb = (torch.randn(2, 4, 32, 32, generator=g) * 0.7).to(dev)
    known = (torch.randn(2, 4, 32, 32, generator=g) * 0.18215).to(dev)
    mask = (torch.from_numpy(np.random.RandomState(7).rand(32, 32) < 0.5).float())[None, None].to(dev)
    s = B200Sampler(unet, DDIMNoiseScheduler(**SCHED_KW), self_condition=True)
    noise = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(1))
    s.sample(rgb, 6, noise=noise, mask=mask, known_latents=known, ddpm=True, seed=1)           # builds + captures
    st = s._state[(2, 32, 6, True, True, False, 0)]
    assert st["graph"] is not None
    n0 = nat.launch_count()
    a = s.sample(rgb, 6, noise=noise, mask=mask, known_latents=known, ddpm=True, seed=1)
    per_call = nat.launch_count() - n0
    assert per_call < 60, per_call          # graph replays: only the input staging kernels are launched from the host
    b = s.sample(rgb, 6, noise=noise, mask=mask, known_latents=known, ddpm=True, seed=1)
    c = s.sample(rgb, 6, noise=noise, mask=mask, known_latents=known, ddpm=True, seed=2)
    d = s.sample(rgb, 6, noise=noise, mask=mask, known_latents=known, ddpm=True, seed=2, ddpm_noise="device")
    assert rel_l2(b, a) <= 5e-3 and rel_l2(c, a) > 5e-2 and torch.isfinite(d).all()
    torch.testing.assert_close(a * mask, known * mask, rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------------------------ §8(f)1
def test_decode_panoptic_vs_compute_pq_transcription(dev, vaes):
    """§8(f)1: decode + per-image post-processing on the device vs the transcription of
    trainers_ldm_cond.py:1243-1313 applied to the oracle's logits of the SAME latents.  Image sizes are those of
    COCO val examples (portrait / landscape / padded)."""
    from oracle import ldmseg_restated as orc
    _, ref_vs, _, vs = vaes
    g = torch.Generator().manual_seed(17)
    z = torch.randn(3, 4, 64, 64, generator=g) * 0.9
    sizes = [(427, 640), (640, 480), (333, 500)]
    crops = [(0, 0, 342, 512), (0, 0, 512, 384), (0, 0, 341, 512)]      # CropResize keeps the aspect ratio, pads the rest
    with torch.no_grad():
        logits = orc.decode_latents(z, ref_vs)
    pads = []
    for (y0, x0, ch, cw) in crops:
        pm = torch.zeros(512, 512)
        pm[y0:y0 + ch, x0:x0 + cw] = 1
        pads.append(pm)
    # a random-init decoder's 128 logits are nearly flat (max probability just above 1 / 128): the released thresholds
    # (0.5 / 512 / 0.5) would void everything, so the rules are exercised with thresholds scaled to that regime; the
    # released values are covered by tools/kernel_check.py --group panoptic on structured logits
    th = dict(mask_th=0.015, count_th=900, overlap_th=0.005)      # oracle: ~30 of 128 segments survive, ~70 % void
    ref = orc.panoptic_postprocess(logits, sizes, th["mask_th"], th["count_th"], th["overlap_th"], 0, True, padding_masks=pads)
    # decode_latents scales by 1 / scaling_factor before decoding (trainers_ldm_cond.py:421)
    out = vs.decode_panoptic(z.to(dev) / vs.scaling_factor, sizes, crops, **th)
    agrees = []
    for i, ((ids, segs), (rid, rsegs)) in enumerate(zip(out, ref)):
        assert tuple(ids.shape) == sizes[i] and ids.dtype == torch.uint8
        agrees.append(float((ids.numpy() == rid).mean()))
    record("decode_panoptic", id_agreement=agrees, segments_gpu=[len(o[1]) for o in out],
           segments_oracle=[len(r[1]) for r in ref], void_fraction=[float((o[0] == 0).float().mean()) for o in out])
    # the flat logits put ~2 % of the pixels within the decoder's bf16 error of mask_th, and a segment whose area sits
    # at count_th flips as a whole: agreement is a loose bar here, exactness is kernel_check's job
    assert min(agrees) >= 0.90 and max(len(r[1]) for r in ref) >= 10


# ------------------------------------------------------------------------------------------------ §8(f)2
def test_generate_stream_equals_generate(dev, unets, vaes):
    """§8(f)2: the overlapped input pipeline (next batch's H2D + VAE encode on a side stream under the graph loop)
    gives the ids of the plain per-batch calls."""
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle.make_golden import SCHED_KW
    _, unet = unets
    _, _, vi, vs = vaes
    s = B200Sampler(unet, DDIMNoiseScheduler(**SCHED_KW), vi, vs, self_condition=True)
    g = torch.Generator().manual_seed(8)
    batches = [torch.rand(1, 3, 256, 256, generator=g).pin_memory() for _ in range(3)]
    ref = [s.generate(b.to(dev), 4, seed=42)[0].cpu() for b in batches]
    got = [ids.cpu() for ids, _ in s.generate_stream(batches, 4, seed=42)]
    assert len(got) == 3
    agree = [float((a == b).float().mean()) for a, b in zip(got, ref)]
    record("generate_stream", id_agreement=agree)
    assert min(agree) >= 0.98          # the side-stream encode never splits K: rounding-level differences only


# ------------------------------------------------------------------------------------------------ §8(f)3
def test_cross_attention_unet_and_guidance(dev):
    """§8(f)3: the stock SD-v1 UNet WITH cross-attention (full width, 8-channel conv_in) on the CUDA engine:
    one forward with text-shaped encoder_hidden_states [2,77,768] at a 32x32 latent (rel-L2 <= 2e-2 vs the fp32
    oracle) and the classifier-free-guidance loop of sample() (multiplier 2, guidance 7.5, 6 steps; <= 1e-1: the
    guidance combine amplifies the per-forward error by ~2 g)."""
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.models import UNet
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle import ldmseg_restated as orc
    from oracle.make_golden import SCHED_KW
    torch.manual_seed(0)
    o = orc.UNet()
    o.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="copy", cond_channels=0)
    o = o.eval()
    unet = UNet()
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="copy", cond_channels=0)
    unet.load_state_dict(o.state_dict(), strict=True)
    unet = unet.to(dev)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 8, 32, 32, generator=g)
    enc = torch.randn(2, 77, 768, generator=g)
    tt = torch.tensor(699)
    y = unet(x.to(dev), tt.to(dev), encoder_hidden_states=enc.to(dev)).sample
    with torch.no_grad():
        ref = o(x, tt, encoder_hidden_states=enc).sample
    e_fwd = rel_l2(y, ref)
    rgb = torch.randn(1, 4, 32, 32, generator=g) * 0.7
    with torch.no_grad():
        ref_lat = orc.sample(o, orc.DDIMNoiseScheduler(**SCHED_KW), rgb, 6, seed=7, self_condition=False,
                             encoder_hidden_states=enc, guidance_scale=7.5)
    s = B200Sampler(unet, DDIMNoiseScheduler(**SCHED_KW), self_condition=False)
    lat = s.sample(rgb.to(dev), 6, seed=7, encoder_hidden_states=enc.to(dev), guidance_scale=7.5)
    e_cfg = rel_l2(lat, ref_lat)
    record("cross_attention_and_guidance", forward_rel_l2=e_fwd, cfg_loop6_rel_l2=e_cfg)
    assert e_fwd <= 2e-2 and e_cfg <= 1e-1
    with pytest.raises(RuntimeError):
        unet(x.to(dev), tt.to(dev), encoder_hidden_states=None)          # cross-attention kept: states are required


# ------------------------------------------------------------------------------------------------ §8(f)4
def test_training_step_forward_reuse(dev, unets):
    """§8(f)4: the training step's no-grad self-conditioning forward (trainers_ldm_cond.py:813-831) on the sampling
    kernels: per-sample timesteps, add_noise fused with the UNet-input write, remove_noise."""
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle import ldmseg_restated as orc
    from oracle.make_golden import SCHED_KW
    oracle_unet, unet = unets
    g = torch.Generator().manual_seed(12)
    lat = torch.randn(4, 4, 32, 32, generator=g) * 0.9
    rgb = torch.randn(4, 4, 32, 32, generator=g) * 0.7
    noise = torch.randn(4, 4, 32, 32, generator=g)
    t = torch.tensor([999, 613, 250, 17])
    with torch.no_grad():
        r_noisy, r_pred, r_cond = orc.self_condition_estimate(oracle_unet, orc.DDIMNoiseScheduler(**SCHED_KW), lat, rgb,
                                                              noise, t)
    sch = DDIMNoiseScheduler(**SCHED_KW)
    noisy, pred, cond = unet.self_condition_estimate(sch, lat.to(dev), rgb.to(dev), noise.to(dev), t.to(dev))
    e = dict(noisy=rel_l2(noisy, r_noisy), pred=rel_l2(pred, r_pred), cond=rel_l2(cond, r_cond))
    record("training_step_forward_reuse", **e)
    assert e["noisy"] <= 1e-6 and e["pred"] <= 2e-2 and e["cond"] <= 5e-2
    # the scheduler's own add_noise / remove_noise on CUDA tensors (per-sample timesteps gathered on the device)
    an = sch.add_noise(lat.to(dev), noise.to(dev), t.to(dev))
    torch.testing.assert_close(an.cpu(), r_noisy, rtol=1e-6, atol=1e-6)
    rn = sch.remove_noise(an, noise.to(dev), t.to(dev))
    torch.testing.assert_close(rn.cpu(), lat, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ configs[2] share
def test_batch_independence_at_config3_size(dev, unets):
    """A size-independent property at BASELINE configs[2]'s per-GPU share (batch 8, 64x64 latent): every image's
    chain is independent of its batch mates, although the batch-8 plan uses other kernels than the
    batch-1 plan (CTA pairs, the stream-K tail, two-tile attention CTAs against split-K, single-tile CTAs).  Ten
    steps of the fused sampler on 8 images against the same images one at a time."""
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle.make_golden import SCHED_KW
    _, unet = unets
    sampler = B200Sampler(unet, DDIMNoiseScheduler(**SCHED_KW))
    g = torch.Generator().manual_seed(11)
    rgb = (torch.randn(8, 4, 64, 64, generator=g) * 0.18215 * 4).to(dev)
    noise = torch.randn(8, 4, 64, 64, generator=g)
    z8 = sampler.sample(rgb, 10, seed=0, noise=noise).clone()
    errs = []
    for i in (0, 3, 7):
        z1 = sampler.sample(rgb[i:i + 1].contiguous(), 10, seed=0, noise=noise[i:i + 1].contiguous())
        errs.append(rel_l2(z8[i:i + 1], z1))
    record("batch_independence_b8_vs_b1_10steps", rel_l2=errs)
    assert max(errs) < 5e-3, errs
