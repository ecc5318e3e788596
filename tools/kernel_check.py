"""Per-kernel numerical checks on a real B200 (run under gpurun).  Each group runs in its own
process (a device trap in one kernel must not hide the others):

    python tools/kernel_check.py            # all groups, each in a subprocess with a timeout
    python tools/kernel_check.py --group igemm
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))

GROUPS = ["simple", "igemm_plain", "igemm_conv", "igemm_epi", "igemm_splitk", "igemm_streamk", "igemm_pair", "igemm_bn320", "igemm_s2", "igemm_up2", "igemm_f32stream", "igemm_lnfold",
          "gn_fused", "norm", "attn_simple", "attn", "xattn", "elementwise", "sampler", "panoptic", "vae_pdl"]


def report(name, got, ref, tol):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    rel = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
    bad = not (rel <= tol) or not torch.isfinite(got).all().item()
    print(f"{'FAIL' if bad else 'PASS'} {name:58s} max_abs={err:.4e} ref_max={scale:.3e} rel_l2={rel:.3e}",
          flush=True)
    return not bad


def run_group(group):
    import torch
    import torch.nn.functional as F
    from ldmseg import _native as nat
    from ldmseg import _pack as pk

    torch.manual_seed(0)
    # the references below must be fp32, not TF32 (cudnn.allow_tf32 defaults to True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    ok = True
    bf = torch.bfloat16

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev) * scale).to(bf)

    def conv_case(name, nb, h, w, cin, cout, *, taps=9, bias=True, residual=False, rowbias=False,
                  act=nat.ACT_NONE, out_f32=False, block_n=0, split_k=0, simple=False, extra_src=None,
                  shortcut=False, stats=False, pdl=False, tiled=False, pair=False, stream_k=False, split_cluster=False):
        """out = conv(x (+ extra_src concat)) [+ 1x1 shortcut of the raw sources] ..."""
        nonlocal ok
        srcs_c = [cin] + ([extra_src] if extra_src else [])
        xs = [rnd(nb * h * w, c) for c in srcs_c]
        ctot = sum(srcs_c)
        kscale = 1.0 / (ctot * taps) ** 0.5
        if taps == 9:
            wt = torch.randn(cout, ctot, 3, 3, device=dev) * kscale
            packed = pk.split_conv3x3_k(wt, srcs_c)
        else:
            wt = torch.randn(cout, ctot, device=dev) * kscale
            packed = pk.split_linear_k(wt, srcs_c)
        segs = [(i, taps) for i in range(len(srcs_c))]
        srcs, src_cs = list(xs), list(srcs_c)
        wsc = None
        if shortcut:
            # extra 1x1 segments over two more raw sources
            sc_c = [cin]
            sc_x = [rnd(nb * h * w, c) for c in sc_c]
            wsc = torch.randn(cout, sum(sc_c), device=dev) / sum(sc_c) ** 0.5
            packed = torch.cat([packed, pk.split_linear_k(wsc, sc_c)], dim=1)
            for x_, c_ in zip(sc_x, sc_c):
                srcs.append(x_)
                src_cs.append(c_)
                segs.append((len(srcs) - 1, 1))
        tiled = tiled or pair
        wb = pk.to_bf16(pk.tile_pack(packed)) if tiled else pk.to_bf16(packed)
        b = torch.randn(cout, device=dev) if bias else None
        rb = torch.randn(nb, cout, device=dev) if rowbias else None
        n_out = cout // 2 if act == nat.ACT_GEGLU else cout
        res = rnd(nb * h * w, cout) if residual else None
        out = torch.full((nb * h * w, n_out), float("nan"), device=dev,
                         dtype=torch.float32 if out_f32 else bf)
        ws = cnt = None
        if split_k > 1 or stream_k:
            ws = torch.full((16 * 1024 * 1024,), float("nan"), device=dev)   # partials are overwritten, never read stale
            cnt = torch.zeros(8192, device=dev, dtype=torch.int32)
        st = torch.zeros(nb, cout, 2, device=dev) if stats else None
        p = nat.make_igemm_params(srcs, src_cs, nb, h, w, segs, wb, cout, out, n_out, bias=b,
                                  rowbias=rb, rowbias_ld=cout, residual=res, res_ld=cout, act=act,
                                  block_n=block_n, split_k=split_k, workspace=ws, counters=cnt, stats=st, pdl=pdl,
                                  weight_tiled=tiled, pair=pair, stream_k=stream_k, split_cluster=split_cluster)
        nat.igemm(p, simple=simple)
        if split_k > 1 or stream_k:  # second launch: tile counters must have reset themselves
            if st is not None:
                st.zero_()
            nat.igemm(p, simple=simple)
        torch.cuda.synchronize()
        # reference in fp32 from the bf16-rounded operands
        xcat = torch.cat([x.float() for x in xs], dim=1)
        wq = wt.to(bf).float()
        if taps == 9:
            xi = xcat.reshape(nb, h, w, ctot).permute(0, 3, 1, 2)
            ref = F.conv2d(xi, wq, padding=1).permute(0, 2, 3, 1).reshape(nb * h * w, cout)
        else:
            ref = xcat @ wq.t()
        if shortcut:
            ref = ref + torch.cat([x.float() for x in srcs[len(xs):]], dim=1) @ wsc.to(bf).float().t()
        if bias:
            ref = ref + b
        if rowbias:
            ref = ref + rb.repeat_interleave(h * w, dim=0)
        if act == nat.ACT_GEGLU:
            r = ref.reshape(nb * h * w, cout // 32, 2, 16)
            ref = (r[:, :, 0] * F.gelu(r[:, :, 1])).reshape(nb * h * w, cout // 2)
        if residual:
            ref = ref + res.float()
        if act == nat.ACT_SILU:
            ref = F.silu(ref)
        ok &= report(name, out, ref, 1e-2 if not out_f32 else 5e-3)
        if stats:
            o = out.float().reshape(nb, h * w, cout)
            ok &= report(name + " [stats sum]", st[:, :, 0], o.sum(1), 1e-3)
            ok &= report(name + " [stats sumsq]", st[:, :, 1], (o * o).sum(1), 1e-3)
        if split_k > 1 or stream_k:
            ok &= report(name + " [counters reset]", cnt.float(), torch.zeros_like(cnt).float(), 0.0)

    if group == "simple":
        conv_case("simple linear 256x320->320", 1, 1, 256, 320, 320, taps=1, simple=True)
        conv_case("simple conv3x3 16x16 64->64", 1, 16, 16, 64, 64, simple=True)
        conv_case("simple conv3x3 2x8x8 128->64 +res+rowbias silu", 2, 8, 8, 128, 64, residual=True,
                  rowbias=True, act=nat.ACT_SILU, simple=True)
        conv_case("simple linear geglu 128x64->128", 1, 1, 128, 64, 128, taps=1, act=nat.ACT_GEGLU,
                  simple=True)
        conv_case("simple conv3x3 dual-source + shortcut", 1, 16, 16, 64, 64, extra_src=128,
                  shortcut=True, simple=True)
    elif group == "igemm_plain":
        for bn in (64, 128, 160, 256):
            conv_case(f"igemm linear 4096x320->640 bn={bn}", 1, 1, 4096, 320, 640, taps=1, block_n=bn)
        conv_case("igemm linear 4096x320->320 auto", 1, 1, 4096, 320, 320, taps=1)
        conv_case("igemm linear 64x1280->1280 (M<128)", 1, 1, 64, 1280, 1280, taps=1)
        conv_case("igemm linear 8192x1280->5120 f32 bias", 1, 1, 8192, 1280, 5120, taps=1, out_f32=True)
        conv_case("igemm linear 256x24->8 (K,N tiny)", 1, 1, 256, 24, 8, taps=1)
    elif group == "igemm_conv":
        conv_case("igemm conv3x3 1x64x64 320->320", 1, 64, 64, 320, 320)
        conv_case("igemm conv3x3 2x32x32 640->640", 2, 32, 32, 640, 640)
        conv_case("igemm conv3x3 1x16x16 1280->1280", 1, 16, 16, 1280, 1280)
        conv_case("igemm conv3x3 1x8x8 1280->1280 (M=64)", 1, 8, 8, 1280, 1280)
        conv_case("igemm conv3x3 3x8x8 1280->640 (odd nb)", 3, 8, 8, 1280, 640)
        conv_case("igemm conv3x3 1x64x64 16->320 (conv_in)", 1, 64, 64, 16, 320)
        conv_case("igemm conv3x3 1x64x64 320->4 f32 (conv_out)", 1, 64, 64, 320, 4, out_f32=True)
        conv_case("igemm conv3x3 1x128x128 64->64", 1, 128, 128, 64, 64)
        conv_case("igemm conv3x3 1x256x256 64->128", 1, 256, 256, 64, 128)
    elif group == "igemm_epi":
        conv_case("igemm conv3x3 +res +rowbias", 2, 32, 32, 320, 320, residual=True, rowbias=True)
        conv_case("igemm conv3x3 silu", 1, 32, 32, 128, 256, act=nat.ACT_SILU)
        conv_case("igemm linear geglu 1024x640->5120", 1, 1, 1024, 640, 5120, taps=1, act=nat.ACT_GEGLU)
        conv_case("igemm conv3x3 dual-source 1280+640->1280", 1, 16, 16, 1280, 1280, extra_src=640)
        conv_case("igemm conv3x3 dual + 1x1 shortcut", 1, 32, 32, 640, 640, extra_src=320, shortcut=True,
                  residual=False)
        conv_case("igemm conv3x3 +res +stats 2x32x32", 2, 32, 32, 320, 320, residual=True, stats=True)
        conv_case("igemm conv3x3 +stats 3x8x8 (tile spans images)", 3, 8, 8, 1280, 640, stats=True)
        conv_case("igemm linear +stats pdl 1x64x64", 1, 64, 64, 320, 320, taps=1, stats=True, pdl=True)
        conv_case("simple conv3x3 +stats", 2, 8, 8, 64, 64, stats=True, simple=True)
        conv_case("simple conv3x3 tiled weights n=40", 1, 8, 8, 64, 40, simple=True, tiled=True)
        for bn in (64, 128, 160, 256):
            conv_case(f"igemm conv3x3 tiled weights bn={bn}", 1, 32, 32, 320, 640, block_n=bn, tiled=True, residual=True)
        conv_case("igemm conv_out tiled n=4 f32", 1, 64, 64, 320, 4, out_f32=True, tiled=True)
        conv_case("igemm tiled split-K dual + shortcut", 1, 16, 16, 1280, 1280, extra_src=640, shortcut=True, split_k=5,
                  tiled=True, stats=True)
    elif group == "gn_fused":
        # GroupNorm whose statistics come from the producers' epilogues (two sources, plain-GEMM producer)
        for (nb, hh, c0, c1, silu) in [(2, 16, 640, 320, True), (1, 64, 320, 0, True), (1, 8, 1280, 1280, True),
                                       (3, 32, 128, 0, False), (2, 8, 1280, 640, True), (1, 32, 640, 640, False),
                                       (1, 64, 640, 320, True)]:
            hw = hh * hh
            outs, sts = [], []
            for c in (c0, c1):
                if c == 0:
                    outs.append(None)
                    sts.append(None)
                    continue
                x = rnd(nb * hw, c)
                wt = torch.randn(c, c, device=dev) / c ** 0.5
                o = torch.empty(nb * hw, c, device=dev, dtype=bf)
                st = torch.zeros(nb, c, 2, device=dev)
                p = nat.make_igemm_params([x], [c], 1, 1, nb * hw, [(0, 1)], pk.to_bf16(pk.pack_linear(wt)), c, o, c,
                                          stats=st, stats_hw=hw)
                nat.igemm(p)
                outs.append(o)
                sts.append(st)
            C = c0 + c1
            g = torch.randn(C, device=dev)
            be = torch.randn(C, device=dev)
            y = torch.empty(nb * hw, C, device=dev, dtype=bf)
            nat.groupnorm_apply_cs(outs[0], c0, sts[0], outs[1], c1, sts[1], nb, hw, 32, g, be, 1e-5, silu, y)
            torch.cuda.synchronize()
            xc = torch.cat([o.float() for o in outs if o is not None], dim=1).reshape(nb, hw, C).permute(0, 2, 1)
            ref = F.group_norm(xc, 32, g, be, 1e-5)
            ref = (F.silu(ref) if silu else ref).permute(0, 2, 1).reshape(nb * hw, C)
            ok &= report(f"groupnorm from fused producer statistics nb={nb} hw={hw} c={c0}+{c1} silu={silu}", y, ref, 1e-2)
    elif group == "igemm_splitk":
        conv_case("igemm conv3x3 1x8x8 1280->1280 split4", 1, 8, 8, 1280, 1280, split_k=4)
        conv_case("igemm conv3x3 1x16x16 1280->1280 split8 +res", 1, 16, 16, 1280, 1280, split_k=8,
                  residual=True)
        conv_case("igemm linear 256x1280->1280 split3 silu", 1, 1, 256, 1280, 1280, taps=1, split_k=3,
                  act=nat.ACT_SILU)
        conv_case("igemm conv3x3 1x8x8 1280->1280 split14 bn64 +stats", 1, 8, 8, 1280, 1280, split_k=14,
                  block_n=64, stats=True, rowbias=True)
        conv_case("igemm linear geglu 256x1280->10240 split2", 1, 1, 256, 1280, 10240, taps=1, split_k=2,
                  act=nat.ACT_GEGLU)
        # the same exchange through distributed shared memory (the splits of a tile = one cluster)
        for s_ in (2, 3, 4, 5, 6, 7, 8, 9, 12, 14, 16):
            print(f"max clusters of {s_}: bn64 {nat.max_split_clusters(64, False, s_)} bn128 "
                  f"{nat.max_split_clusters(128, False, s_)} bn160 {nat.max_split_clusters(160, False, s_)} bn256 "
                  f"{nat.max_split_clusters(256, False, s_)} geglu bn256 {nat.max_split_clusters(256, True, s_)}")
        conv_case("cluster split conv3x3 1x8x8 1280->1280 split4 bn128", 1, 8, 8, 1280, 1280, split_k=4, block_n=128,
                  split_cluster=True, tiled=True)
        conv_case("cluster split conv3x3 1x16x16 1280->1280 split9 bn160 +res", 1, 16, 16, 1280, 1280, split_k=9,
                  block_n=160, residual=True, split_cluster=True, tiled=True)
        conv_case("cluster split linear 256x1280->1280 split3 bn256 silu", 1, 1, 256, 1280, 1280, taps=1, split_k=3,
                  block_n=256, act=nat.ACT_SILU, split_cluster=True, tiled=True)
        conv_case("cluster split conv3x3 1x8x8 1280->1280 split6 bn64 +stats +rowbias", 1, 8, 8, 1280, 1280,
                  split_k=6, block_n=64, stats=True, rowbias=True, split_cluster=True, tiled=True)
        conv_case("cluster split conv3x3 1x64x64 320->320 split2 bn160 pdl", 1, 64, 64, 320, 320, split_k=2,
                  block_n=160, pdl=True, split_cluster=True, tiled=True)
        conv_case("cluster split conv3x3 1x32x32 640->640 split5 bn256 f32", 1, 32, 32, 640, 640, split_k=5,
                  block_n=256, out_f32=True, split_cluster=True, tiled=True)
        conv_case("cluster split linear geglu 256x1280->10240 split2 bn256", 1, 1, 256, 1280, 10240, taps=1,
                  split_k=2, block_n=256, act=nat.ACT_GEGLU, split_cluster=True, tiled=True)
        conv_case("cluster split conv3x3 1x8x8 1280->1280 split14 bn128 (falls back unless 10 clusters of 14 fit)",
                  1, 8, 8, 1280, 1280, split_k=14, block_n=128, split_cluster=True, tiled=True)
        conv_case("cluster split conv3x3 dual + 1x1 shortcut 1x16x16 split8 bn160", 1, 16, 16, 1280, 640,
                  extra_src=640, shortcut=True, split_k=8, block_n=160, split_cluster=True, tiled=True)
    elif group == "igemm_streamk":
        # stream-K tail: ragged last waves cut along K (tiles mod 148, or mod 74 pairs)
        conv_case("streamk conv3x3 8x64x64 320->320 bn160 (512 tiles: 3 waves + 68) +res +rowbias +stats", 8, 64, 64,
                  320, 320, block_n=160, residual=True, rowbias=True, stats=True, tiled=True, stream_k=True)
        conv_case("streamk pair conv3x3 8x64x64 320->320 bn160 (256 pair tiles: 3 waves + 34) +stats", 8, 64, 64,
                  320, 320, block_n=160, stats=True, pair=True, stream_k=True)
        conv_case("streamk conv3x3 8x32x32 640->640 bn160 (256 tiles: 1 wave + 108) silu", 8, 32, 32, 640, 640,
                  block_n=160, act=nat.ACT_SILU, tiled=True, stream_k=True)
        conv_case("streamk pair conv3x3 8x32x32 640->640 bn256 (96 pair tiles: 1 wave + 22) pdl", 8, 32, 32, 640,
                  640, block_n=256, pair=True, pdl=True, stream_k=True)
        conv_case("streamk conv3x3 8x16x16 1280->1280 bn256 (80 tiles, all tail) f32", 8, 16, 16, 1280, 1280,
                  block_n=256, out_f32=True, tiled=True, stream_k=True)
        conv_case("streamk conv3x3 1x64x64 320->320 bn160 (64 tiles, all tail)", 1, 64, 64, 320, 320, block_n=160,
                  tiled=True, stream_k=True)
        conv_case("streamk linear 8192x1280->1280 bn128 (640 tiles: 4 waves + 48)", 1, 1, 8192, 1280, 1280, taps=1,
                  block_n=128, tiled=True, stream_k=True)
        conv_case("streamk pair linear 4224x1280->640 bn128 (odd m tiles, phantom)", 1, 1, 4224, 1280, 640, taps=1,
                  block_n=128, pair=True, stream_k=True)
        conv_case("streamk conv3x3 dual + 1x1 shortcut 8x32x32 bn160", 8, 32, 32, 640, 640, extra_src=320,
                  shortcut=True, block_n=160, tiled=True, stream_k=True)
        conv_case("streamk full last wave falls back to whole tiles (296 tiles)", 1, 1, 9472, 640, 512, taps=1,
                  block_n=128, tiled=True, stream_k=True)
    elif group == "igemm_pair":
        # CTA pairs (cta_group::2): 256 x block_n tiles
        for bn in (128, 160, 256):
            conv_case(f"pair linear 4096x320->640 bn={bn}", 1, 1, 4096, 320, 640, taps=1, block_n=bn, pair=True)
        conv_case("pair conv3x3 1x64x64 320->320 bn160 +res +rowbias +stats", 1, 64, 64, 320, 320, block_n=160,
                  residual=True, rowbias=True, stats=True, pair=True)
        conv_case("pair conv3x3 2x32x32 640->640 bn256 pdl", 2, 32, 32, 640, 640, block_n=256, pair=True, pdl=True)
        conv_case("pair linear 384x640->1280 bn256 (odd tile count)", 1, 1, 384, 640, 1280, taps=1, block_n=256,
                  pair=True, stats=True)
        conv_case("pair linear geglu 1024x640->5120 bn256", 1, 1, 1024, 640, 5120, taps=1, act=nat.ACT_GEGLU,
                  block_n=256, pair=True)
        conv_case("pair linear 8192x1280->5120 f32 (9 waves)", 1, 1, 8192, 1280, 5120, taps=1, out_f32=True,
                  block_n=256, pair=True)
        conv_case("pair conv3x3 8x64x64 320->320 bn160 (4 waves) +stats", 8, 64, 64, 320, 320, block_n=160,
                  stats=True, pair=True)
        conv_case("pair conv3x3 dual + 1x1 shortcut bn160", 1, 32, 32, 640, 640, extra_src=320, shortcut=True,
                  block_n=160, pair=True)
        conv_case("pair conv3x3 1x16x16 1280->1280 bn256 split4 +res", 1, 16, 16, 1280, 1280, block_n=256,
                  split_k=4, residual=True, pair=True)
        conv_case("pair linear 384x1280->1280 bn160 split3 (odd tiles) +stats", 1, 1, 384, 1280, 1280, taps=1,
                  block_n=160, split_k=3, stats=True, pair=True)
        conv_case("pair conv3x3 1x64x64 320->4 f32 bn128 (conv_out)", 1, 64, 64, 320, 4, out_f32=True, block_n=128,
                  pair=True)
    elif group == "igemm_bn320":
        # 320-wide pair tiles: two N = 160 tcgen05.mma per k-step over three accumulator slots (ring), the first k-blocks'
        # first-half MMAs issued ahead of the second-half ones
        conv_case("bn320 pair linear 4096x320->640 (1 tile per pair)", 1, 1, 4096, 320, 640, taps=1, block_n=320, pair=True)
        conv_case("bn320 pair linear 256x128->320 (k shorter than the run-ahead)", 1, 1, 256, 128, 320, taps=1,
                  block_n=320, pair=True)
        conv_case("bn320 pair conv3x3 1x64x64 320->320 +res +rowbias +stats", 1, 64, 64, 320, 320, block_n=320,
                  residual=True, rowbias=True, stats=True, pair=True)
        conv_case("bn320 pair conv3x3 8x64x64 320->320 (128 pair tiles: 2 waves) +res +rowbias +stats pdl", 8, 64, 64,
                  320, 320, block_n=320, residual=True, rowbias=True, stats=True, pair=True, pdl=True)
        conv_case("bn320 pair linear 40960x320->640 (320 pair tiles: 5 waves: every slot order) silu", 1, 1, 40960, 320,
                  640, taps=1, block_n=320, pair=True, act=nat.ACT_SILU)
        conv_case("bn320 pair conv3x3 8x32x32 640->640 (64 pair tiles) f32 +stats", 8, 32, 32, 640, 640, block_n=320,
                  out_f32=True, stats=True, pair=True)
        conv_case("bn320 pair conv3x3 dual + 1x1 shortcut 2x32x32 -> 640", 2, 32, 32, 640, 640, extra_src=320,
                  shortcut=True, block_n=320, pair=True)
        conv_case("bn320 pair linear 384x640->1280 (odd tile count, phantom) +stats", 1, 1, 384, 640, 1280, taps=1,
                  block_n=320, pair=True, stats=True)
        conv_case("bn320 pair linear 1024x640->400 (ragged N: second half partly past N)", 1, 1, 1024, 640, 400, taps=1,
                  block_n=320, pair=True)
        conv_case("bn320 streamk pair conv3x3 8x64x64 320->320 (128 pair tiles: 1 wave + 54) +res +stats", 8, 64, 64,
                  320, 320, block_n=320, residual=True, stats=True, pair=True, stream_k=True)
        conv_case("bn320 streamk pair conv3x3 8x16x16 1280->1280 (32 pair tiles, all tail) +rowbias", 8, 16, 16, 1280,
                  1280, block_n=320, rowbias=True, pair=True, stream_k=True)
        conv_case("bn320 streamk pair conv3x3 8x32x32 960->640 dual (64 pair tiles, all tail) pdl", 8, 32, 32, 640, 640,
                  extra_src=320, block_n=320, pair=True, stream_k=True, pdl=True)
    elif group == "norm":
        for (nb, hw, c0, c1) in [(1, 4096, 320, 0), (2, 1024, 640, 320), (1, 256, 1280, 640),
                                 (2, 64, 1280, 1280), (1, 65536, 256, 0)]:
            C = c0 + c1
            x0 = rnd(nb * hw, c0) + 0.5
            x1 = rnd(nb * hw, c1, scale=2.0) if c1 else None
            g = torch.randn(C, device=dev)
            be = torch.randn(C, device=dev)
            out = torch.empty(nb * hw, C, device=dev, dtype=bf)
            stats = torch.empty(nb * 32 * 2, device=dev)
            nat.groupnorm(x0, c0, x1, c1, nb, hw, 32, g, be, 1e-5, True, out, stats)
            torch.cuda.synchronize()
            xc = torch.cat([x0.float()] + ([x1.float()] if c1 else []), dim=1)
            xi = xc.reshape(nb, hw, C).permute(0, 2, 1)
            ref = F.silu(F.group_norm(xi, 32, g, be, 1e-5)).permute(0, 2, 1).reshape(nb * hw, C)
            ok &= report(f"groupnorm+silu nb={nb} hw={hw} c={c0}+{c1}", out, ref, 1e-2)
        for (rows, c) in [(4096, 320), (1024, 640), (300, 1280), (1000, 256)]:
            x = rnd(rows, c) + 0.3
            g = torch.randn(c, device=dev)
            be = torch.randn(c, device=dev)
            out = torch.empty(rows, c, device=dev, dtype=bf)
            nat.layernorm(x, rows, c, g, be, 1e-5, False, out)
            torch.cuda.synchronize()
            ok &= report(f"layernorm rows={rows} c={c}", out, F.layer_norm(x.float(), (c,), g, be, 1e-5),
                         1e-2)
    elif group in ("attn_simple", "attn"):
        simple = group == "attn_simple"
        cases = [(1, 256, 8, 40), (2, 64, 8, 160), (1, 256, 8, 160), (1, 1024, 8, 80),
                 (1, 4096, 8, 40), (2, 1024, 4, 40), (1, 200, 8, 40), (1, 1000, 2, 80), (3, 320, 8, 80),
                 (1, 4096 + 96, 1, 40)]   # ragged token counts: masked key columns, partial query tiles
        for (nb, ntok, heads, d) in cases:
            C = heads * d
            qkv = rnd(nb * ntok, 3 * C, scale=1.5)
            out = torch.full((nb * ntok, C), float("nan"), device=dev, dtype=bf)
            nat.attention(qkv, nb, ntok, heads, d, out, simple=simple)
            torch.cuda.synchronize()
            q, k, v = qkv.float().reshape(nb, ntok, 3, heads, d).permute(2, 0, 3, 1, 4)
            ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(nb * ntok, C)
            ok &= report(f"{group} nb={nb} ntok={ntok} heads={heads} d={d}", out, ref, 2e-2)
    elif group == "elementwise":
        x = rnd(1000, 2 * 640)
        out = torch.empty(1000, 640, device=dev, dtype=bf)
        nat.geglu(x, 1000, 640, out)
        ref = x.float()[:, :640] * F.gelu(x.float()[:, 640:])
        ok &= report("geglu", out, ref, 1e-2)
        x = rnd(2 * 16 * 16, 64)
        out = torch.empty(2 * 32 * 32, 64, device=dev, dtype=bf)
        nat.upsample2x(x, 2, 16, 16, 64, out)
        ref = F.interpolate(x.float().reshape(2, 16, 16, 64).permute(0, 3, 1, 2), scale_factor=2.0,
                            mode="nearest").permute(0, 2, 3, 1).reshape(-1, 64)
        ok &= report("upsample2x", out, ref, 0.0)
        for pad_lo in (1, 0):
            x = rnd(2 * 16 * 16, 64)
            out = torch.empty(2 * 8 * 8, 9 * 64, device=dev, dtype=bf)
            nat.im2col_s2(x, 2, 16, 16, 64, pad_lo, out)
            xi = x.float().reshape(2, 16, 16, 64).permute(0, 3, 1, 2)
            xi = F.pad(xi, (1, 1, 1, 1)) if pad_lo == 1 else F.pad(xi, (0, 1, 0, 1))
            cols = F.unfold(xi, 3, stride=2)  # [nb, C*9, L] with C-major, tap-minor
            L = cols.shape[-1]
            ref = cols.reshape(2, 64, 9, L).permute(0, 3, 2, 1).reshape(2 * L, 9 * 64)
            ok &= report(f"im2col_s2 pad_lo={pad_lo}", out, ref, 0.0)
        src = torch.randn(2, 4, 64 * 64, device=dev)
        out = torch.zeros(2 * 64 * 64, 16, device=dev, dtype=bf)
        nat.nchw_to_nhwc_bf16(src, 2, 4, 64 * 64, 16, 4, 2.0, -1.0, out)
        ref = torch.zeros(2 * 64 * 64, 16, device=dev)
        ref[:, 4:8] = (src * 2 - 1).permute(0, 2, 1).reshape(-1, 4)
        ok &= report("nchw_to_nhwc_bf16", out, ref.to(bf), 0.0)
        src = torch.randn(2 * 4096, 4, device=dev)
        out = torch.empty(2, 4, 4096, device=dev)
        nat.nhwc_f32_to_nchw(src, 2, 4, 4096, 4, 0.5, out)
        ok &= report("nhwc_f32_to_nchw", out, src.reshape(2, 4096, 4).permute(0, 2, 1) * 0.5, 1e-7)
        # ddim step (all prediction types)
        for ptype in (0, 1, 2):
            mo = torch.randn(2, 4, 64, 64, device=dev)
            xs = torch.randn(2, 4, 64, 64, device=dev)
            prev = torch.empty_like(mo)
            x0 = torch.empty_like(mo)
            a_t, a_p = 0.0059128, 0.0046601
            nat.ddim_step(mo, xs, a_t, a_p, ptype, False, 1.0, False, prev, x0)
            at = torch.tensor(a_t)
            ap = torch.tensor(a_p)
            bt = 1 - at
            if ptype == 0:
                r0 = (xs - bt ** 0.5 * mo) / at ** 0.5
                eps = mo
            elif ptype == 1:
                r0 = mo
                eps = (xs - at ** 0.5 * r0) / bt ** 0.5
            else:
                r0 = at ** 0.5 * xs - bt ** 0.5 * mo
                eps = at ** 0.5 * mo + bt ** 0.5 * xs
            rp = ap ** 0.5 * r0 + (1 - ap) ** 0.5 * eps
            ok &= report(f"ddim_step ptype={ptype} prev", prev, rp, 1e-6)
            ok &= report(f"ddim_step ptype={ptype} x0", x0, r0, 1e-6)
        # time embedding
        t = torch.tensor([999.0, 19.0, 500.0], device=dev)
        out = torch.empty(3, 320, device=dev)
        nat.timestep_sinusoid(t, 3, 320, True, 0.0, out)
        half = 160
        fr = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, device=dev) / half)
        a = t[:, None] * fr[None]
        ok &= report("timestep_sinusoid", out, torch.cat([a.cos(), a.sin()], -1), 1e-4)
        x = torch.randn(3, 320, device=dev)
        w = torch.randn(1280, 320, device=dev) / 18
        b = torch.randn(1280, device=dev)
        out = torch.empty(3, 1280, device=dev)
        nat.small_linear(x, 3, 320, w, b, 1280, False, True, out, 1280)
        ok &= report("small_linear silu_out", out, F.silu(x @ w.t() + b), 1e-5)
        # convT shuffle + LN2d + SiLU
        nb, h, w_, c = 1, 16, 16, 256
        src = rnd(nb * h * w_, 4 * c)
        g = torch.randn(c, device=dev)
        be = torch.randn(c, device=dev)
        out = torch.empty(nb * 4 * h * w_, c, device=dev, dtype=bf)
        nat.convt_shuffle_ln(src, nb, h, w_, c, g, be, 1e-6, True, out)
        s = src.float().reshape(nb, h, w_, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(nb, 2 * h, 2 * w_, c)
        ref = F.silu(F.layer_norm(s, (c,), g, be, 1e-6)).reshape(-1, c)
        ok &= report("convt_shuffle_ln", out, ref, 1e-2)
        # bilinear x2
        nb, h, w_, c = 2, 32, 64, 128
        src = rnd(nb * h * w_, c)
        out = torch.empty(nb, c, 2 * h, 2 * w_, device=dev)
        nat.bilinear2x_to_nchw(src, nb, h, w_, c, c, out)
        xi = src.float().reshape(nb, h, w_, c).permute(0, 3, 1, 2)
        ref = F.interpolate(xi, scale_factor=2, mode="bilinear", align_corners=False)
        ok &= report("bilinear2x_to_nchw", out, ref, 1e-5)
        ids = torch.empty(nb, 2 * h, 2 * w_, device=dev, dtype=torch.uint8)
        mp = torch.empty(nb, 2 * h, 2 * w_, device=dev)
        nat.bilinear2x_argmax(src, nb, h, w_, c, c, ids, mp)
        agree = (ids.long() == ref.argmax(1)).float().mean().item()
        print(f"{'PASS' if agree > 0.999 else 'FAIL'} bilinear2x_argmax agreement={agree:.5f}")
        ok &= agree > 0.999
        ok &= report("bilinear2x_argmax maxprob", mp, ref.softmax(1).max(1)[0], 1e-4)

    elif group == "igemm_s2":
        # 3x3 stride-2 convolutions through the TMA traversal stride (no im2col): UNet Downsample2D (padding 1)
        # and the AutoencoderKL encoder (F.pad(0,1,0,1), padding 0); (nb, hout) output geometry
        for (nb, hout, cin, cout, pad, kw) in [
                (1, 32, 320, 320, 1, {}), (2, 16, 640, 640, 1, {}), (1, 8, 1280, 1280, 1, dict(split_k=4)),
                (1, 4, 1280, 1280, 1, {}), (1, 256, 128, 128, 0, {}), (1, 128, 256, 256, 0, {}),
                (2, 64, 512, 512, 0, {}), (1, 256, 32, 64, 1, dict(act=nat.ACT_SILU)), (1, 32, 320, 320, 1, dict(pair=True, block_n=160)),
                (3, 8, 64, 64, 0, dict(simple=True))]:
            hin = 2 * hout
            x = rnd(nb * hin * hin, cin)
            wt = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
            b = torch.randn(cout, device=dev)
            tiled = True
            wb = pk.to_bf16(pk.tile_pack(pk.pack_conv3x3(wt)))
            out = torch.full((nb * hout * hout, cout), float("nan"), device=dev, dtype=bf)
            split = kw.get("split_k", 0)
            ws = torch.full((16 * 1024 * 1024,), float("nan"), device=dev) if split > 1 else None
            cnt = torch.zeros(8192, device=dev, dtype=torch.int32) if split > 1 else None
            p = nat.make_igemm_params([x], [cin], nb, hout, hout, [(0, 9)], wb, cout, out, cout, bias=b,
                                      act=kw.get("act", nat.ACT_NONE), block_n=kw.get("block_n", 0), split_k=split,
                                      workspace=ws, counters=cnt, weight_tiled=tiled, pair=kw.get("pair", False),
                                      weight_static=True, pdl=True, conv_stride=2, conv_pad=pad)
            nat.igemm(p, simple=kw.get("simple", False))
            torch.cuda.synchronize()
            xi = x.float().reshape(nb, hin, hin, cin).permute(0, 3, 1, 2)
            wq = wt.to(bf).float()
            if pad == 1:
                ref = F.conv2d(xi, wq, b, stride=2, padding=1)
            else:
                ref = F.conv2d(F.pad(xi, (0, 1, 0, 1)), wq, b, stride=2, padding=0)
            if kw.get("act") == nat.ACT_SILU:
                ref = F.silu(ref)
            ref = ref.permute(0, 2, 3, 1).reshape(nb * hout * hout, cout)
            ok &= report(f"igemm conv3x3 stride 2 (TMA) {nb}x{hin}->{hout} {cin}->{cout} pad={pad} {kw}", out, ref, 1e-2)
    elif group == "igemm_up2":
        # nearest x2 up-sampling folded into the 3x3 convolution after it (diffusers Upsample2D): four 2x2 phase GEMMs
        # over the input pixels against F.interpolate(nearest) + F.conv2d in fp32; (nb, hin) INPUT geometry.  The UNet's
        # three shapes at batch 1 / 2 / 8 in every schedule the planner may pick, statistics, f32 + shadow outputs
        for (nb, hin, cin, cout, kw) in [
                (1, 8, 1280, 1280, {}), (1, 8, 1280, 1280, dict(split_k=7, block_n=128)),
                (1, 16, 1280, 1280, dict(split_k=3, block_n=256, stats=True)),
                (1, 32, 640, 640, dict(block_n=160, stats=True)), (2, 32, 640, 640, dict(pair=True, block_n=256)),
                (1, 32, 640, 640, dict(pair=True, block_n=160, split_k=2, stats=True)),
                (8, 16, 1280, 1280, dict(pair=True, block_n=256, stream_k=True, stats=True)),
                (8, 32, 640, 640, dict(block_n=160, stream_k=True)), (2, 16, 192, 96, dict(act=nat.ACT_SILU, block_n=64)),
                (2, 8, 64, 72, dict(out_f32=True, stats=True)), (3, 16, 320, 320, dict(pair=True, block_n=128, stream_k=True)),
                (8, 8, 1280, 1280, dict(pair=True, block_n=256, stats=True)),
                (8, 32, 640, 640, dict(pair=True, block_n=320, stats=True)),
                (4, 16, 1280, 1280, dict(pair=True, block_n=320, stream_k=True))]:
            x = rnd(nb * hin * hin, cin)
            wt = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
            b = torch.randn(cout, device=dev)
            wb = pk.to_bf16(pk.tile_pack(pk.pack_upsample2_conv3x3(wt)))
            hout = 2 * hin
            out_f32 = kw.get("out_f32", False)
            out = torch.full((nb * hout * hout, cout), float("nan"), device=dev, dtype=torch.float32 if out_f32 else bf)
            out2 = torch.full((nb * hout * hout, cout), float("nan"), device=dev, dtype=bf) if out_f32 else None
            split, tail = kw.get("split_k", 0), kw.get("stream_k", False)
            ws = torch.full((16 * 1024 * 1024,), float("nan"), device=dev) if (split > 1 or tail) else None
            cnt = torch.zeros(8192, device=dev, dtype=torch.int32) if (split > 1 or tail) else None
            st = torch.zeros(nb, cout, 2, device=dev) if kw.get("stats") else None
            p = nat.make_igemm_params([x], [cin], nb, hin, hin, [(0, 4)], wb, cout, out, cout, bias=b,
                                      act=kw.get("act", nat.ACT_NONE), block_n=kw.get("block_n", 0), split_k=split,
                                      workspace=ws, counters=cnt, weight_tiled=True, pair=kw.get("pair", False),
                                      weight_static=True, pdl=True, stats=st, stats_hw=hin * hin, out2=out2,
                                      stream_k=tail, upsample2=True)
            for _ in range(2):          # twice: split-K / tail counters must have re-armed themselves
                if st is not None:
                    st.zero_()
                nat.igemm(p)
            torch.cuda.synchronize()
            xi = x.float().reshape(nb, hin, hin, cin).permute(0, 3, 1, 2)
            ref = F.conv2d(F.interpolate(xi, scale_factor=2.0, mode="nearest"), wt.to(bf).float(), b, padding=1)
            if kw.get("act") == nat.ACT_SILU:
                ref = F.silu(ref)
            ref = ref.permute(0, 2, 3, 1).reshape(nb * hout * hout, cout)
            name = f"igemm upsample2 + conv3x3 {nb}x{hin}->{hout} {cin}->{cout} {kw}"
            # (a phase weight is the f32 sum of up to four taps rounded to bf16 once: same error class as the taps')
            ok &= report(name, out, ref, 1e-2 if not out_f32 else 6e-3)
            if out2 is not None:
                ok &= report(name + " [bf16 shadow]", out2, ref, 1e-2)
            if st is not None:
                o = out.float().reshape(nb, hout * hout, cout)
                ok &= report(name + " [stats sum]", st[:, :, 0], o.sum(1), 2e-3)
                ok &= report(name + " [stats sumsq]", st[:, :, 1], (o * o).sum(1), 2e-3)
            if cnt is not None:
                ok &= report(name + " [counters reset]", cnt.float(), torch.zeros_like(cnt).float(), 0.0)
    elif group == "igemm_f32stream":
        # fp32 residual stream: f32 residual in, f32 out + bf16 shadow, statistics of the f32 values
        for (nb, h, cin, cout, taps, kw) in [(2, 32, 320, 320, 9, {}), (1, 64, 320, 320, 1, {}),
                                             (1, 16, 1280, 1280, 9, dict(split_k=6)), (1, 64, 320, 4, 9, {}),
                                             (2, 8, 64, 64, 9, dict(simple=True))]:
            m = nb * h * h
            x = rnd(m, cin)
            if taps == 9:
                wt = torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5
                wb = pk.to_bf16(pk.tile_pack(pk.pack_conv3x3(wt)))
            else:
                wt = torch.randn(cout, cin, device=dev) / cin ** 0.5
                wb = pk.to_bf16(pk.tile_pack(pk.pack_linear(wt)))
            b = torch.randn(cout, device=dev)
            res = torch.randn(m, cout, device=dev)
            out = torch.full((m, cout), float("nan"), device=dev)
            out2 = torch.full((m, cout), float("nan"), device=dev, dtype=bf)
            st = torch.zeros(nb, cout, 2, device=dev)
            split = kw.get("split_k", 0)
            ws = torch.full((16 * 1024 * 1024,), float("nan"), device=dev) if split > 1 else None
            cnt = torch.zeros(8192, device=dev, dtype=torch.int32) if split > 1 else None
            p = nat.make_igemm_params([x], [cin], nb, h, h, [(0, taps)], wb, cout, out, cout, bias=b, residual=res,
                                      res_ld=cout, split_k=split, workspace=ws, counters=cnt, stats=st,
                                      weight_tiled=True, weight_static=True, out2=out2)
            nat.igemm(p, simple=kw.get("simple", False))
            torch.cuda.synchronize()
            xi = x.float().reshape(nb, h, h, cin).permute(0, 3, 1, 2)
            wq = wt.to(bf).float()
            if taps == 9:
                ref = F.conv2d(xi, wq, padding=1).permute(0, 2, 3, 1).reshape(m, cout)
            else:
                ref = x.float() @ wq.t()
            ref = ref + b + res
            ok &= report(f"igemm f32 stream out {nb}x{h}x{h} {cin}->{cout} taps={taps} {kw}", out, ref, 5e-3)
            ok &= report("   bf16 shadow", out2, out.to(bf), 0.0)
            if not kw.get("simple"):
                o = out.reshape(nb, h * h, cout)
                ok &= report("   statistics of the f32 values [sum]", st[:, :, 0], o.sum(1), 1e-3)
                ok &= report("   statistics of the f32 values [sumsq]", st[:, :, 1], (o * o).sum(1), 1e-3)
        # GroupNorm / LayerNorm reading the f32 stream
        nb, hw, c0, c1 = 2, 256, 640, 320
        xs = [torch.randn(nb * hw, c, device=dev) * 1.5 + 0.3 for c in (c0, c1)]
        sts = [torch.stack([x.reshape(nb, hw, -1).sum(1), (x * x).reshape(nb, hw, -1).sum(1)], dim=-1).contiguous() for x in xs]
        g, be = torch.randn(c0 + c1, device=dev), torch.randn(c0 + c1, device=dev)
        y = torch.empty(nb * hw, c0 + c1, device=dev, dtype=bf)
        nat.groupnorm_apply_cs(xs[0], c0, sts[0], xs[1], c1, sts[1], nb, hw, 32, g, be, 1e-5, True, y)
        xc = torch.cat(xs, dim=1).reshape(nb, hw, c0 + c1).permute(0, 2, 1)
        ref = F.silu(F.group_norm(xc, 32, g, be, 1e-5)).permute(0, 2, 1).reshape(nb * hw, c0 + c1)
        ok &= report("groupnorm apply, f32 sources", y, ref, 1e-2)
        for (rows, c) in [(4096, 320), (300, 1280)]:
            x = torch.randn(rows, c, device=dev) + 0.3
            g, be = torch.randn(c, device=dev), torch.randn(c, device=dev)
            out = torch.empty(rows, c, device=dev, dtype=bf)
            nat.layernorm(x, rows, c, g, be, 1e-5, False, out)
            ok &= report(f"layernorm f32 source rows={rows} c={c}", out, F.layer_norm(x, (c,), g, be, 1e-5), 1e-2)
    elif group == "igemm_lnfold":
        # LayerNorm folded into the GEMMs around it: the producer accumulates per-row moments of its bf16 output, the
        # consumer multiplies the raw rows by W diag(gamma) and finishes rstd * (acc - mean * colsum) + (W beta + b)
        for (m, c, n, kw) in [(4096, 320, 960, {}), (1024, 640, 5120, dict(geglu=True)), (256, 1280, 10240, dict(geglu=True)),
                              (64, 1280, 3840, {}), (256, 1280, 3840, dict(split_k=3)), (4096, 320, 2560, dict(geglu=True, pair=True, block_n=256)),
                              (300, 320, 960, dict(simple=True)), (8192, 320, 960, dict(f32_stream=True))]:
            x = rnd(m, c)
            wp = torch.randn(c, c, device=dev) / c ** 0.5
            res = rnd(m, c) * 2 + 0.7
            f32s = kw.get("f32_stream", False)
            t = torch.full((m, c), float("nan"), device=dev, dtype=torch.float32 if f32s else bf)
            t_sh = torch.full((m, c), float("nan"), device=dev, dtype=bf) if f32s else None
            rs = torch.zeros(m, 2, device=dev)
            resid = res.float() if f32s else res
            p1 = nat.make_igemm_params([x], [c], 1, 1, m, [(0, 1)], pk.to_bf16(pk.tile_pack(pk.pack_linear(wp))), c, t, c,
                                       residual=resid, res_ld=c, weight_tiled=True, weight_static=True, rowstats_out=rs,
                                       out2=t_sh)
            nat.igemm(p1, simple=kw.get("simple", False))
            trow = t_sh if f32s else t                      # the bf16 rows the consumer multiplies
            gam, bet = torch.randn(c, device=dev) * 0.5 + 1.0, torch.randn(c, device=dev) * 0.3
            w = torch.randn(n, c, device=dev) / c ** 0.5
            b = torch.randn(n, device=dev)
            wf, cf, colsum = pk.fold_layernorm(w, b, gam, bet)
            geglu = kw.get("geglu", False)
            if geglu:
                wf, cf = pk.interleave_geglu(wf, cf)
                colsum = wf.to(bf).float().sum(dim=1)
            n_out = n // 2 if geglu else n
            out = torch.full((m, n_out), float("nan"), device=dev, dtype=bf)
            split = kw.get("split_k", 0)
            ws = torch.full((16 * 1024 * 1024,), float("nan"), device=dev) if split > 1 else None
            cnt = torch.zeros(8192, device=dev, dtype=torch.int32) if split > 1 else None
            p2 = nat.make_igemm_params([trow], [c], 1, 1, m, [(0, 1)], pk.to_bf16(pk.tile_pack(pk.pack_linear(wf))), n, out, n_out,
                                       bias=cf.contiguous(), act=nat.ACT_GEGLU if geglu else nat.ACT_NONE, weight_tiled=True,
                                       weight_static=True, split_k=split, workspace=ws, counters=cnt,
                                       pair=kw.get("pair", False), block_n=kw.get("block_n", 0), pdl=True,
                                       ln_rowstats=rs, ln_colsum=colsum.contiguous(), ln_channels=c, ln_eps=1e-5)
            nat.igemm(p2, simple=kw.get("simple", False))
            torch.cuda.synchronize()
            tr = trow.float()
            # the moments are taken from the f32 values before the bf16 rounding of the stored row
            ok &= report(f"ln-fold producer row moments m={m} c={c} {kw}", rs, torch.stack([tr.sum(1), (tr * tr).sum(1)], 1), 1e-3)
            ref = F.layer_norm(tr, (c,), gam, bet, 1e-5) @ w.t() + b
            if geglu:
                ref = ref[:, : n // 2] * F.gelu(ref[:, n // 2:])
            ok &= report(f"ln-fold consumer m={m} c={c} n={n} {kw}", out, ref, 1.2e-2)
    elif group == "xattn":
        # cross-attention: q [nb*ntok, C], kv [nb*T, 2C]; T = 77 (text) / 257 (CLIP patches) / 1 / 128 (queries)
        for (nb, ntok, T, heads, d) in [(2, 4096, 77, 8, 40), (2, 1024, 257, 8, 80), (2, 256, 77, 8, 160),
                                        (2, 64, 77, 8, 160), (1, 1024, 1, 8, 80), (2, 4096, 128, 8, 40),
                                        (4, 256, 257, 8, 160)]:
            C = heads * d
            q = rnd(nb * ntok, C, scale=1.5)
            kv = rnd(nb * T, 2 * C, scale=1.5)
            out = torch.full((nb * ntok, C), float("nan"), device=dev, dtype=bf)
            nat.cross_attention(q, kv, nb, ntok, T, heads, d, out)
            torch.cuda.synchronize()
            qq = q.float().reshape(nb, ntok, heads, d).permute(0, 2, 1, 3)
            kk, vv = kv.float().reshape(nb, T, 2, heads, d).permute(2, 0, 3, 1, 4)
            ref = F.scaled_dot_product_attention(qq, kk, vv).permute(0, 2, 1, 3).reshape(nb * ntok, C)
            ok &= report(f"cross-attention nb={nb} ntok={ntok} T={T} heads={heads} d={d}", out, ref, 2e-2)
    elif group == "sampler":
        # fused sampler step against the term-by-term formula, all prediction types / clip / guidance / extensions
        m, n = 2 * 64 * 64, 7
        coef = torch.rand(n, 4, device=dev) * 0.8 + 0.1
        sig = torch.rand(n, device=dev) * 0.3
        for (ptype, clip, cfg, selfc, inpaint, ddpm, step) in [(0, False, False, True, False, False, 3),
                                                                (1, False, False, True, False, False, 0),
                                                                (2, True, False, True, False, False, 6),
                                                                (0, True, True, False, False, False, 2),
                                                                (0, False, False, True, True, True, 4),
                                                                (0, False, False, True, True, True, 6)]:
            eps = torch.randn((2 if cfg else 1) * m, 4, device=dev)
            lat = torch.randn(m, 4, device=dev)
            lat0 = lat.clone()
            x0 = torch.empty(m, 4, device=dev)
            rgb = torch.randn(m, 4, device=dev)
            xin = torch.full(((2 if cfg else 1) * m, 16), float("nan"), device=dev, dtype=bf)
            stp = torch.tensor([step], device=dev, dtype=torch.int32)
            mask = (torch.rand(m, device=dev) < 0.5).float() if inpaint else None
            known = torch.randn(n, m, 4, device=dev) if inpaint else None
            noise = torch.randn(n, m, 4, device=dev) if ddpm else None
            gs = 7.5
            nat.sampler_step(eps, lat, x0, rgb, xin, m, coef, stp, n, selfc, mask, known, noise, sig if ddpm else None,
                             ptype, clip, 0.8, cfg, gs)
            torch.cuda.synchronize()
            e = eps[:m] + gs * (eps[m:] - eps[:m]) if cfg else eps
            sa, sb, pa, pb = coef[step]
            if ptype == 0:
                r0, ee = (lat0 - sb * e) / sa, e
            elif ptype == 1:
                r0, ee = e, (lat0 - sa * e) / sb
            else:
                r0, ee = sa * lat0 - sb * e, sa * e + sb * lat0
            if clip:
                r0 = r0.clamp(-0.8, 0.8)
            pv = pa * r0 + pb * ee
            last = step == n - 1
            if ddpm and not last:
                pv = pv + sig[step] * noise[step]
            nxt = r0 if last else pv
            if inpaint:
                nxt = mask[:, None] * known[step] + (1 - mask[:, None]) * nxt
            tag = f"sampler_step ptype={ptype} clip={clip} cfg={cfg} inpaint={inpaint} ddpm={ddpm} step={step}"
            ok &= report(tag + " latents", lat, nxt, 2e-6)
            ok &= report(tag + " x0", x0, r0, 2e-6)
            row = torch.cat([lat, rgb, x0 if selfc else torch.zeros_like(x0), torch.zeros(m, 4, device=dev)], 1).to(bf)  # the kernel's own f32 results, rounded
            ok &= report(tag + " unet_in", xin[:m], row, 0.0)
            if cfg:
                ok &= report(tag + " unet_in (second half)", xin[m:], row, 0.0)
        # select_row / advance_step
        table = torch.randn(9, 20160, device=dev)
        stp = torch.tensor([4], device=dev, dtype=torch.int32)
        dst = torch.empty(3, 20160, device=dev)
        nat.select_row(table, 20160, stp, 3, dst)
        nat.advance_step(stp)
        ok &= report("select_row", dst, table[4:5].expand(3, -1), 0.0)
        ok &= report("advance_step", stp.float(), torch.tensor([5.0], device=dev), 0.0)
        # softmax_rows (f32 scores -> bf16 probabilities)
        for (rows, cols) in [(512, 4096), (128, 16384), (300, 1024)]:
            sc = torch.randn(rows, cols, device=dev) * 3
            pr = torch.empty(rows, cols, device=dev, dtype=bf)
            nat.softmax_rows(sc, rows, cols, 0.044, pr)
            ok &= report(f"softmax_rows {rows}x{cols}", pr, torch.softmax(sc * 0.044, dim=-1), 1e-2)
        # add_noise / remove_noise with per-sample device timesteps (+ fused UNet-input write)
        acp = torch.cumprod(1 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, device=dev) ** 2, 0)
        nb, hw = 3, 32 * 32
        x = torch.randn(nb, 4, 32, 32, device=dev)
        nz = torch.randn(nb, 4, 32, 32, device=dev)
        t = torch.tensor([999, 19, 500], device=dev)
        out = torch.empty_like(x)
        xin = torch.zeros(nb * hw, 16, device=dev, dtype=bf)
        nat.noise_mix(x, nz, t, acp, nb, 4 * hw, 1.0, 0, out, xin, hw, 16)
        a = acp[t].view(-1, 1, 1, 1)
        ref = a ** 0.5 * x + (1 - a) ** 0.5 * nz
        ok &= report("noise_mix add_noise", out, ref, 1e-6)
        ok &= report("noise_mix add_noise -> unet_in", xin[:, :4], out.permute(0, 2, 3, 1).reshape(-1, 4).to(bf), 0.0)
        back = torch.empty_like(x)
        nat.noise_mix(out, nz, t, acp, nb, 4 * hw, 1.0, 1, back)
        ok &= report("noise_mix remove_noise", back, (ref - (1 - a) ** 0.5 * nz) / a ** 0.5, 1e-5)
    elif group == "panoptic":
        # decode tail + post-processing against the transcription of compute_pq (oracle): float stage by agreement,
        # integer stage (area rules + relabel) bit-exact
        sys.path.insert(0, ROOT)
        from oracle import ldmseg_restated as orc
        import numpy as np
        g = torch.Generator().manual_seed(3)
        nb, s = 3, 64
        # one confident winner per 8x8 cell over a suppressed background, so that every rule of the filter fires:
        # classes 1..9 large segments (kept), class 0 = ignore label (dropped), class 10 a single small cell (dropped by
        # count_th), class 11 wins two cells but is weakly positive over many more (dropped by overlap_th)
        win = torch.randint(0, 10, (nb, 8, 8), generator=g)
        win[:, 0, 0] = 10
        win[:, 7, 6:8] = 11
        base = torch.full((nb, 128, 8, 8), -8.0)
        base.scatter_(1, win[:, None], 8.0)
        halo = (win != 11) & (torch.rand(nb, 8, 8, generator=g) < 0.6)
        base[:, 11][halo] = 1.5
        logits = F.interpolate(base, size=(s, s), mode="nearest") + torch.randn(nb, 128, s, s, generator=g) * 0.3
        sizes = [(100, 150), (128, 128), (97, 61)]
        crops = [(0, 0, 2 * s, 2 * s), (0, 0, 2 * s, 100), (8, 4, 96, 120)]
        cl = logits.permute(0, 2, 3, 1).contiguous().to(dev)
        geom = torch.tensor([[h, w, *c] for (h, w), c in zip(sizes, crops)], dtype=torch.int32, device=dev)
        max_hw = max(h * w for h, w in sizes)
        stride = (max_hw + 15) // 16 * 16
        pred = torch.full((nb, stride), -7, device=dev, dtype=torch.int16)
        area = torch.empty(nb, 128, device=dev, dtype=torch.int32)
        orig = torch.empty(nb, 128, device=dev, dtype=torch.int32)
        ids = torch.zeros(nb, stride, device=dev, dtype=torch.uint8)
        keep = torch.empty(nb, 128, device=dev, dtype=torch.int32)
        count_th, overlap_th = 300, 0.5
        nat.panoptic_resample(cl, nb, s, 128, 128, geom, max_hw, stride, 0.5, True, pred, area, orig)
        nat.panoptic_filter(pred, nb, geom, max_hw, stride, area, orig, count_th, overlap_th, 0, ids, keep)
        torch.cuda.synchronize()
        up = F.interpolate(logits, scale_factor=2, mode="bilinear", align_corners=False)      # vae.py:270
        pads = []
        for (y0, x0, ch, cw) in crops:
            pm = torch.zeros(2 * s, 2 * s)
            pm[y0:y0 + ch, x0:x0 + cw] = 1
            pads.append(pm)
        ref = orc.panoptic_postprocess(up, sizes, 0.5, count_th, overlap_th, 0, True, padding_masks=pads)
        for i, (h, w) in enumerate(sizes):
            got = ids[i, :h * w].view(h, w).cpu().numpy()
            agree = (got == ref[i][0]).mean()
            kept = sorted(int(c) + 1 for c in keep[i].nonzero().flatten().cpu())
            good = agree >= 0.995 and kept == sorted(ref[i][1])
            good = good and len(kept) >= 2
            print(f"{'PASS' if good else 'FAIL'} panoptic image {i} ({h}x{w}): id agreement {agree:.5f}, segments {len(kept)} vs {len(ref[i][1])}, "
                  f"void {float((got == 0).mean()):.3f}")
            ok &= bool(good)
            # integer stage: bit-exact given the kernel's own pred / histograms
            p_np = pred[i, :h * w].view(h, w).cpu().numpy().astype(np.int64)
            a_np, o_np = area[i].cpu().numpy(), orig[i].cpu().numpy()
            assert (np.bincount(p_np[p_np >= 0].ravel(), minlength=128) == a_np).all(), "area histogram"
            exp_ids, exp_keep = orc.panoptic_filter(p_np, a_np, o_np, None, count_th, overlap_th, 0)
            exact = (got == exp_ids).all() and kept == exp_keep
            print(f"{'PASS' if exact else 'FAIL'} panoptic image {i}: integer stage bit-exact")
            ok &= bool(exact)
    elif group == "vae_pdl":
        # the AutoencoderKL mid-block attention feeds tensors produced on the stream as igemm "weights": under PDL
        # they must not be fetched before the grid-dependency wait.  Encoder at 128 / 256 px, PDL on (many
        # repetitions) vs PDL off.
        from ldmseg.models import GeneralVAEImage
        from ldmseg.engine import plan as plan_mod
        torch.manual_seed(0)
        vae = GeneralVAEImage().to(dev)
        for size, nb in ((128, 2), (256, 2)):
            # two different batches alternate, so a launch that read an operand before its producer had written it
            # would see the OTHER batch's data (an O(1) error), not a harmless copy of the right values
            xs = [torch.rand(nb, 3, size, size, device=dev) * 2 - 1 for _ in range(2)]
            eng = vae._get_engine()
            pl = eng.plan(nb, size)
            pl.pdl = False
            refs = [eng.encode(x).clone() for x in xs]
            spread = max(((eng.encode(x) - r).norm() / r.norm()).item() for x, r in zip(xs, refs))
            pl.pdl = True
            worst = 0.0
            for it in range(20):
                out = eng.encode(xs[it & 1])
                worst = max(worst, ((out - refs[it & 1]).norm() / refs[it & 1].norm()).item())
            other = ((refs[0] - refs[1]).norm() / refs[1].norm()).item()
            # run-to-run spread without PDL (fp32 atomics order -> bf16 rounding flips, amplified by the peaked
            # single-head softmax of the random-init mid block) is the yardstick
            good = worst < max(3e-2, 3 * spread) and other > 0.3
            print(f"{'PASS' if good else 'FAIL'} VAE encoder {size}px nb={nb}: PDL vs no-PDL worst rel_l2 over 20 alternating runs = "
                  f"{worst:.3e} (no-PDL run-to-run spread {spread:.3e}, distance between the two batches {other:.3e})")
            ok &= good
    torch.cuda.synchronize()
    print(f"GROUP {group}: {'OK' if ok else 'FAILED'}", flush=True)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.group:
        sys.exit(0 if run_group(args.group) else 1)
    bad = []
    for g in GROUPS:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g],
                               timeout=args.timeout)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -999
            print(f"GROUP {g}: TIMEOUT", flush=True)
        print(f"--- group {g} rc={rc} ({time.time() - t0:.1f}s)", flush=True)
        if rc != 0:
            bad.append(g)
    print("FAILED GROUPS:", bad if bad else "none")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
