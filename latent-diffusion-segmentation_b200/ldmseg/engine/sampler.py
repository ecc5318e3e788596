"""The whole denoising loop on the device: a CUDA-graph replay per step, no host synchronisation.

Semantics are those of `TrainerDiffusion.sample`
(/root/reference/ldmseg/trainers/trainers_ldm_cond.py:1045-1170) in its released configuration
(no descriptor model, multiplier = 1, optional self-conditioning):

    noise ~ torch.Generator().manual_seed(seed) ON THE CPU, copied to the GPU          (:1088-1092)
    for t in scheduler.timesteps:  eps = unet(cat[latents, rgb, x0_prev], t)           (:1127-1141)
        x0, x_prev = ddim_step(eps, t, latents);  condition = x0                       (:1150-1159)
        latents = x0 on the last step, else x_prev                                     (:1154-1159)

What is different from driving `unet(...)` + `scheduler.step(...)` from Python:
  * the time-embedding MLP and all 22 `time_emb_proj` are evaluated ONCE for all timesteps (t is
    batch-uniform and known in advance) -> a [steps, 20160] table; each step selects its row;
  * one fused kernel does the scheduler update, writes the fp32 latent state, x0 and the next
    step's 16-channel bf16 UNet input (no torch.cat, no dtype cast, no .item());
  * the step index lives in device memory, so ONE captured graph (267 kernel nodes) is replayed
    N times.
Extensions that the reference lacks, both exact restatements of the oracle
in oracle/ldmseg_restated.py: sampling-time inpainting (`mask`, `known_latents`) and ancestral
DDPM noise (`ddpm=True`).
"""
from __future__ import annotations

from typing import Optional

import torch

from ldmseg import _native as nat


class B200Sampler:
    def __init__(self, unet, scheduler, vae_image=None, vae_semseg=None, self_condition: bool = True,
                 use_graph: bool = True):
        self.unet, self.scheduler = unet, scheduler
        self.vae_image, self.vae_semseg = vae_image, vae_semseg
        self.self_condition = self_condition
        self.use_graph = use_graph
        self._state = {}

    # ------------------------------------------------------------------ setup per (B, L, steps)
    def _prepare(self, nb: int, size: int, steps: int, ddpm: bool):
        key = (nb, size, steps, ddpm)
        st = self._state.get(key)
        if st is not None:
            return st
        eng = self.unet._get_engine()
        dev = eng.device
        sch = self.scheduler
        sch.set_timesteps_inference(steps)
        ts = sch.timesteps.cpu()
        n = len(ts)
        ratio = sch.num_train_timesteps // sch.num_inference_steps
        acp = sch.alphas_cumprod.float()
        coef = torch.zeros(n, 4)
        sigma = torch.zeros(n)
        for i, t in enumerate(ts.tolist()):
            a_t = acp[t]
            a_p = acp[t - ratio] if t - ratio >= 0 else sch.final_alpha_cumprod.float()
            if ddpm and i != n - 1:
                var = ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)).clamp(min=0)
                sigma[i] = var ** 0.5
            coef[i] = torch.stack([a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5,
                                   (1 - a_p - sigma[i] ** 2).clamp(min=0) ** 0.5])
        with torch.cuda.device(dev):
            plan = eng.plan(nb, size)
            m = nb * size * size
            st = dict(
                plan=plan, n=n, m=m, timesteps=ts,
                coef=coef.to(dev), sigma=sigma.to(dev) if ddpm else None,
                temb_all=torch.empty(n, eng.weights.temb_total, device=dev),
                step=torch.zeros(1, device=dev, dtype=torch.int32),
                lat=torch.zeros(m, 4, device=dev), x0=torch.zeros(m, 4, device=dev),
                rgb=torch.zeros(m, 4, device=dev), graph=None, graph_key=None,
            )
            eng.weights.time_embedding(ts.to(dev).float().contiguous(), st["temb_all"])
        self._state[key] = st
        return st

    def _step_body(self, st, mask, known, noise):
        plan, w = st["plan"], st["plan"].W
        nat.select_row(st["temb_all"], w.temb_total, st["step"], plan.nb, plan.temb)
        plan.run()
        nat.sampler_step(plan.eps, st["lat"], st["x0"], st["rgb"], plan.x_in, st["m"], st["coef"], st["step"],
                         st["n"], self.self_condition, mask, known, noise, st["sigma"])
        nat.advance_step(st["step"])

    @property
    def launches_per_step(self) -> int:
        st = next(iter(self._state.values()))
        return st["plan"].n_launch + 3

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def sample(self, rgb_latents: torch.Tensor, num_inference_steps: int = 50, seed: Optional[int] = None,
               noise: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
               known_latents: Optional[torch.Tensor] = None, ddpm: bool = False) -> torch.Tensor:
        """rgb_latents [B,4,L,L] on the GPU -> final latents f32 [B,4,L,L] (x0 of the last step)."""
        if not rgb_latents.is_cuda:
            raise RuntimeError("B200Sampler.sample needs CUDA tensors (no CPU fallback)")
        nb, _, size, _ = rgb_latents.shape
        st = self._prepare(nb, size, num_inference_steps, ddpm)
        plan, dev, m, n = st["plan"], st["plan"].device, st["m"], st["n"]
        hw = size * size
        with torch.cuda.device(dev):
            if noise is None:
                gen = torch.Generator().manual_seed(seed) if seed is not None else None
                noise = torch.randn((nb, 4, size, size), generator=gen)          # CPU draw, as the reference
            lat0 = (noise.to(dev, non_blocking=True).float() * self.scheduler.init_noise_sigma).contiguous()
            rgb = rgb_latents.float().contiguous()
            nat.nchw_f32_to_nhwc(lat0, nb, 4, hw, 4, 1.0, st["lat"])
            nat.nchw_f32_to_nhwc(rgb, nb, 4, hw, 4, 1.0, st["rgb"])
            plan.x_in.zero_()                                                    # condition = zeros (:1126)
            nat.nchw_to_nhwc_bf16(lat0, nb, 4, hw, plan.W.cin_pad, 0, 1.0, 0.0, plan.x_in)
            nat.nchw_to_nhwc_bf16(rgb, nb, 4, hw, plan.W.cin_pad, 4, 1.0, 0.0, plan.x_in)
            st["step"].zero_()
            mk = kn = nz = None
            if mask is not None:
                # extension: known region re-noised to the next level with the initial noise
                mk = mask.to(dev).float().expand(nb, 1, size, size).reshape(nb, hw).reshape(-1).contiguous()
                ts = st["timesteps"]
                kl = []
                for i in range(n):
                    if i == n - 1:
                        k = known_latents.to(dev).float()
                    else:
                        k = self.scheduler.add_noise(known_latents.to(dev).float(), lat0, ts[i + 1].expand(nb))
                    kl.append(k.permute(0, 2, 3, 1).reshape(m, 4))
                kn = torch.stack(kl).contiguous()
            if ddpm:
                g2 = torch.Generator().manual_seed(1234)
                z = torch.stack([torch.randn((nb, 4, size, size), generator=g2) for _ in range(n - 1)] +
                                [torch.zeros(nb, 4, size, size)])
                nz = z.to(dev).permute(0, 1, 3, 4, 2).reshape(n, m, 4).contiguous()
            gkey = (mk is not None, nz is not None)
            if self.use_graph and mk is None and nz is None:
                if st["graph"] is None:
                    # warm-up step outside capture (lazy kernel attribute setup), then restore state
                    keep = (st["lat"].clone(), plan.x_in.clone())
                    self._step_body(st, None, None, None)
                    torch.cuda.synchronize()
                    st["lat"].copy_(keep[0])
                    plan.x_in.copy_(keep[1])
                    st["step"].zero_()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._step_body(st, None, None, None)
                    st["graph"] = g
                    st["lat"].copy_(keep[0])
                    plan.x_in.copy_(keep[1])
                    st["step"].zero_()
                for _ in range(n):
                    st["graph"].replay()
            else:
                for _ in range(n):
                    self._step_body(st, mk, kn, nz)
            out = torch.empty(nb, 4, size, size, device=dev)
            nat.nhwc_f32_to_nchw(st["lat"], nb, 4, hw, 4, 1.0, out)
        return out

    # ------------------------------------------------------------------ end-to-end: RGB -> panoptic ids
    @torch.no_grad()
    def encode_rgb(self, images: torch.Tensor) -> torch.Tensor:
        """images f32 [B,3,S,S] in [0,1] -> rgb latents (encode_inputs, trainers_ldm_cond.py:334-394):
        2x-1 fused into the layout conversion, posterior mode, x scaling_factor."""
        moments = self.vae_image._get_engine().encode(images, in_scale=2.0, in_shift=-1.0)
        return moments[:, :4] * self.vae_image.scaling_factor

    @torch.no_grad()
    def generate(self, images: torch.Tensor, num_inference_steps: int = 50, seed: Optional[int] = 42):
        """RGB batch on the GPU -> (panoptic ids u8 [B,S,S], max-prob f32 [B,S,S])."""
        rgb_latents = self.encode_rgb(images)
        latents = self.sample(rgb_latents, num_inference_steps, seed=seed)
        # decode_latents: latents * (1 / scaling_factor) then decode (trainers_ldm_cond.py:421-422)
        return self.vae_semseg._get_engine().decode_ids(latents, scale=1.0 / self.vae_semseg.scaling_factor)
