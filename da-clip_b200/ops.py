"""Host-side operator layer: weight repacking into the kernels' native layouts and typed wrappers over the
C ABI.  Repacking runs once per checkpoint load (torch is used here as memory + layout plumbing only)."""
import ctypes as C
import os

import torch

from . import lib as L

TILE_SHAPES = [(1, 128), (2, 64), (4, 32), (8, 16), (16, 8)]


def choose_tile(oh, ow, ndy=1):
    """Pixel-tile (h, w) with h*w == 128 that loads the fewest activation rows: tiles * (h + ndy - 1) * w, where
    ndy vertically adjacent taps share one load (ties: widest)."""
    best, best_cost = None, None
    for th, tw in TILE_SHAPES:
        cost = (-(-oh // th)) * (-(-ow // tw)) * (th + ndy - 1) * tw
        if best is None or cost < best_cost:
            best, best_cost = (th, tw), cost
    return best


def choose_block_n(cout):
    """UMMA N and padded channel count for a plain epilogue."""
    if cout <= 256:
        bn = -(-cout // 16) * 16
        return bn, bn
    for bn in (256, 128, 64):
        if cout % bn == 0:
            return bn, cout
    bn = 128
    return bn, -(-cout // bn) * bn


def _pad_rows(w, cout_pad):
    """w: [Z, cout, cin] -> zero-padded [Z, cout_pad, cin] bf16 contiguous."""
    z, cout, cin = w.shape
    out = torch.zeros(z, cout_pad, cin, dtype=torch.bfloat16, device=w.device)
    out[:, :cout] = w.to(torch.bfloat16)
    return out.contiguous()


class PackedWeight:
    """Weights in kernel layout bf16 [Z][cout_pad][cin] plus the tap geometry that indexes Z."""

    def __init__(self, w, taps, cout, stride=1, ngroups=1, out_scale=1, out_off=((0, 0),), cols=None):
        self.w = w                      # [Z, cout_pad, cin] bf16
        self.taps = taps                # per group: list of (dy, dx)
        # per group: column groups (dx, dy0, [tap indices of dy0, dy0+1, ...]) - vertically adjacent taps that can
        # share one activation load; None = one load per tap
        self.cols = cols
        self.cout = cout
        self.stride = stride
        self.ngroups = ngroups
        self.out_scale = out_scale
        self.out_off = out_off          # per group (oy, ox)


def pack_conv(weight, cout_pad=None, stride=1, pad=None):
    """nn.Conv2d weight [cout, cin, kh, kw] -> taps-major K-major slabs.  Tap (ky,kx) reads input
    (y*stride + ky - pad, x*stride + kx - pad)."""
    cout, cin, kh, kw = weight.shape
    pad = kh // 2 if pad is None else pad
    cout_pad = cout_pad or choose_block_n(cout)[1]
    w = weight.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    taps = [[(ky - pad, kx - pad) for ky in range(kh) for kx in range(kw)]]
    cols = [[(kx - pad, -pad, [ky * kw + kx for ky in range(kh)]) for kx in range(kw)]] if stride == 1 and kh > 1 else None
    return PackedWeight(_pad_rows(w, cout_pad), taps, cout, stride=stride, cols=cols)


def pack_linear(weight, cout_pad=None):
    """nn.Linear weight [out, in] (or a 1x1 conv weight [out, in, 1, 1])."""
    w = weight.reshape(weight.shape[0], weight.shape[1])
    cout = w.shape[0]
    cout_pad = cout_pad or choose_block_n(cout)[1]
    return PackedWeight(_pad_rows(w[None], cout_pad), [[(0, 0)]], cout)


def pack_geglu(weight, bias, block_n=256):
    """GEGLU proj (attention.py:37-44): rows [0,inner) are values, [inner,2*inner) gates.  Rows are
    interleaved per N tile as [block_n/2 values | block_n/2 matching gates] so that one accumulator tile holds
    both halves of the same channels.  Returns (PackedWeight, permuted bias)."""
    two_inner, cin = weight.shape
    inner, half = two_inner // 2, block_n // 2
    assert inner % half == 0
    idx = []
    for t in range(inner // half):
        idx += list(range(t * half, (t + 1) * half)) + list(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx, device=weight.device)
    pw = PackedWeight(_pad_rows(weight[idx][None], two_inner), [[(0, 0)]], two_inner)
    return pw, bias[idx].float().contiguous()


def pack_stem(weight):
    """init_conv 7x7, cin=6 (arch.py:36): K = (kx, c) packed into 64 channels (8 per kx), 7 vertical taps;
    pairs with dac_unet_stem_input."""
    cout, cin, kh, kw = weight.shape
    assert (cin, kh, kw) == (6, 7, 7)
    w = torch.zeros(kh, cout, 64, dtype=weight.dtype, device=weight.device)
    for kx in range(kw):
        w[:, :, kx * 8:kx * 8 + cin] = weight[:, :, :, kx].permute(2, 0, 1)
    taps = [[(ky - 3, 0) for ky in range(kh)]]
    return PackedWeight(_pad_rows(w, choose_block_n(cout)[1]), taps, cout, cols=[[(0, -3, list(range(kh)))]])


def pack_stem_pair(weight):
    """init_conv in pixel-pair form (dac_unet_stem_input with pair=1): A row = the 8-wide window (kx' = 0..7, 8 channels
    each) around a pair of horizontally adjacent pixels; 2 * cout weight rows - pixel 2j's copy at kx = kx', pixel 2j+1's
    at kx = kx' - 1 - so the accumulator row holds both outputs ([B, Hp, Wp/2, 2 cout] == [B, Hp, Wp, cout])."""
    cout, cin, kh, kw = weight.shape
    assert (cin, kh, kw) == (6, 7, 7) and 2 * cout <= 256
    w = torch.zeros(kh, 2 * cout, 64, dtype=weight.dtype, device=weight.device)
    for kx in range(kw):
        w[:, :cout, kx * 8:kx * 8 + cin] = weight[:, :, :, kx].permute(2, 0, 1)
        w[:, cout:, (kx + 1) * 8:(kx + 1) * 8 + cin] = weight[:, :, :, kx].permute(2, 0, 1)
    taps = [[(ky - 3, 0) for ky in range(kh)]]
    return PackedWeight(_pad_rows(w, choose_block_n(2 * cout)[1]), taps, 2 * cout, cols=[[(0, -3, list(range(kh)))]])


def pack_upsample_conv(weight):
    """nearest-2x upsample followed by a 3x3 conv (module_util.py:100-104) == four 2x2 convs on the
    low-resolution input, one per output parity (py, px), whose taps are sums of the original taps that hit the
    same source pixel.  2.25x fewer MACs and no materialised upsampled tensor."""
    cout, cin, kh, kw = weight.shape
    assert (kh, kw) == (3, 3)
    # parity 0: source offsets -1 <- {k0}, 0 <- {k1,k2};  parity 1: 0 <- {k0,k1}, +1 <- {k2}
    sets = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}
    slabs, taps, offs, cols = [], [], [], []
    w32 = weight.float()
    for py in (0, 1):
        for px in (0, 1):
            gt = []
            for dy, kys in sets[py]:
                for dx, kxs in sets[px]:
                    acc = torch.zeros(cout, cin, dtype=torch.float32, device=weight.device)
                    for ky in kys:
                        for kx in kxs:
                            acc += w32[:, :, ky, kx]
                    slabs.append(acc)
                    gt.append((dy, dx))
            taps.append(gt)
            offs.append((py, px))
            # tap index = a*2 + b (a: row offset index, b: column offset index): the two rows of a column share a load
            cols.append([(sets[px][b][0], sets[py][0][0], [b, 2 + b]) for b in (0, 1)])
    w = torch.stack(slabs)                                   # [16, cout, cin]
    cout_pad = choose_block_n(cout)[1]
    return PackedWeight(_pad_rows(w, cout_pad), taps, cout, ngroups=4, out_scale=2, out_off=tuple(offs), cols=cols)


class ConvPlan:
    """One dac_conv plan (TMA descriptors baked for fixed buffers)."""

    def __init__(self, src0, c0, pw: PackedWeight, out=None, *, B, H, W, src1=None, c1=0, ld0=None, ld1=None,
                 epi=L.EPI_PLAIN, act=L.ACT_NONE, bias=None, bias_img=None, film=None, film_off=0,
                 ln_g=None, ln_eps=1e-5, res=None, res2=None, res_f32=None, out_f32=None, out_planar=None, out_coff=0,
                 out_nchw=None,
                 per_image_w=False, block_n=None, weight_override=None, tile=None, share_taps=True,
                 rsrc0=None, rc0=0, rsrc1=None, rc1=0, rweight=None, stats_out=None, stats_eps=1e-5,
                 ln_stats=None, ln_colsum=None, kv_shift=None, ctx_acc=None, halo=None):
        L.require_cuda(src0)
        lib = L.load()
        d = L.ConvDesc()
        s = pw.stride
        OH, OW = (H // s, W // s) if s == 2 else (H, W)
        wt = weight_override if weight_override is not None else pw.w
        cout_pad = wt.shape[-2]
        if block_n is None:
            if cout_pad <= 256:
                block_n = choose_block_n(pw.cout)[0]
            else:
                # several N tiles: pick the tile width whose tile count quantises best onto the 148 persistent CTAs
                # (e.g. 12800 tokens x 768 outputs: 300 tiles of 256 need 3 rounds of 148, 400 tiles of 192 also 3)
                m_tiles = -(-(B * OH * OW) // 128) * pw.ngroups
                cands = [bn for bn in (256, 192, 128) if cout_pad % bn == 0 and (epi == L.EPI_PLAIN or bn != 192)]
                block_n = min(cands or [128], key=lambda bn: (-(-(m_tiles * (cout_pad // bn)) // 148) * (bn + 24), -bn))
        # vertically adjacent taps share one activation load when the weight tiles are small enough to ride along
        share = pw.cols is not None and s == 1 and block_n <= 128 and share_taps
        cols = pw.cols if share else [[(dx, dy, [i]) for i, (dy, dx) in enumerate(gt)] for gt in pw.taps]
        halo_req = halo
        if halo is None:
            # 3x3 stride-1 convs whose whole weight tensor stays resident in shared memory (the 64-/128-channel
            # layers, which are shared-memory-bandwidth bound): ONE haloed activation load per K chunk instead of
            # three column loads - 2.7x less L2 -> SM traffic and shared-memory write traffic
            halo = int(share and len(pw.taps[0]) == 9 and pw.ngroups == 1 and not per_image_w and tile is None
                       and cout_pad == block_n and 9 * ((c0 + c1) // 64) * block_n * 128 <= 150 * 1024
                       and not os.environ.get("DAC_NO_HALO"))
        halo4 = False
        if halo_req is None or halo_req:
            # ... and the folded upsample conv (four parity groups of 2 x 2 taps): one (16+1) x (8+1) box per K chunk and
            # group instead of two column loads - the N = 64 instance at 128^2 -> 256^2 is bound by L2 -> SM ingest
            halo4 = (pw.ngroups == 4 and len(pw.taps[0]) == 4 and s == 1 and share and tile is None and not per_image_w
                     and cout_pad == block_n and 4 * ((c0 + c1) // 64) * block_n * 128 <= 150 * 1024
                     and not os.environ.get("DAC_NO_HALO"))
        if halo4:
            halo = 1
            # taps of a group are stored as (a, b) = (row, column) offset index, a * 2 + b: origin = the smallest offsets
            cols = [[(min(dx for _, dx in gt), min(dy for dy, _ in gt), [0, 1, 2, 3])] for gt in pw.taps]
            tile = (16, 8)
        elif halo:                   # nine shifted operand views of one (16+2) x (8+2) pixel box
            assert len(pw.taps[0]) == 9 and s == 1 and pw.ngroups == 1
            cols, tile = [[(-1, -1, list(range(9)))]], (16, 8)
        ndy = len(cols[0][0][2])
        th, tw = tile or choose_tile(OH, OW, ndy)
        d.src0, d.c0, d.ld0 = src0.data_ptr(), c0, ld0 or src0.shape[-1]
        if src1 is not None:
            d.src1, d.c1, d.ld1 = src1.data_ptr(), c1, ld1 or src1.shape[-1]
        d.B, d.H, d.W, d.OH, d.OW = B, H, W, OH, OW
        d.stride, d.ngroups, d.ntaps = s, pw.ngroups, len(pw.taps[0])
        d.ndy, d.ncols = ndy, len(cols[0])
        for g, gc in enumerate(cols):
            for j, (dx, dy0, tap_ids) in enumerate(gc):
                d.col_dx[g][j], d.col_dy0[g][j] = dx, dy0
                for i, tid in enumerate(tap_ids):
                    d.col_tap[g][j * ndy + i] = tid
        d.out_scale = pw.out_scale
        for g, (oy, ox) in enumerate(pw.out_off):
            d.out_oy[g], d.out_ox[g] = oy, ox
        d.weight, d.cout, d.cout_pad, d.per_image_w = wt.data_ptr(), pw.cout, cout_pad, int(per_image_w)
        d.block_n, d.tile_h, d.tile_w = block_n, th, tw
        if rsrc0 is not None:        # fused 1x1 skip conv (rweight: PackedWeight of the [cout, rc0+rc1] matrix)
            d.rsrc0, d.rc0, d.rld0 = rsrc0.data_ptr(), rc0, rsrc0.shape[-1]
            if rsrc1 is not None:
                d.rsrc1, d.rc1, d.rld1 = rsrc1.data_ptr(), rc1, rsrc1.shape[-1]
            d.rweight = rweight.w.data_ptr()
        d.epi, d.act, d.halo = epi, act, int(halo)
        d.bias = bias.data_ptr() if bias is not None else None
        d.bias_img = bias_img.data_ptr() if bias_img is not None else None
        if film is not None:
            d.film, d.film_ld, d.film_off = film.data_ptr(), film.shape[-1], film_off
        if ln_g is not None:
            d.ln_g, d.ln_eps = ln_g.data_ptr(), ln_eps
        if res is not None:
            d.res, d.res_ld = res.data_ptr(), res.shape[-1]
        if res2 is not None:
            d.res2, d.res2_ld = res2.data_ptr(), res2.shape[-1]
        if out is not None:
            d.out, d.out_ld, d.out_coff = out.data_ptr(), out.shape[-1], out_coff
        if res_f32 is not None:
            d.res_f32, d.res_f32_ld = res_f32.data_ptr(), res_f32.shape[-1]
        if out_f32 is not None:
            d.out_f32, d.out_f32_ld = out_f32.data_ptr(), out_f32.shape[-1]
        if out_planar is not None:
            d.out_planar = out_planar.data_ptr()
        if stats_out is not None:
            d.stats_out, d.stats_eps = stats_out.data_ptr(), stats_eps
        if ln_stats is not None:
            d.ln_stats, d.ln_colsum = ln_stats.data_ptr(), ln_colsum.data_ptr()
        if kv_shift is not None:     # EPI_KVCTX: k | v reduced into the LinearAttention context in the epilogue
            d.kv_shift, d.ctx_acc = kv_shift.data_ptr(), ctx_acc.data_ptr()
            d.ctx_slots = ctx_acc.shape[-2]              # [B, 4, slots, 1088]: see ctx_slots()
        if out_nchw is not None:
            d.out_nchw = out_nchw.data_ptr()
            d.out_nchw_c, d.out_nchw_h, d.out_nchw_w = out_nchw.shape[1], out_nchw.shape[2], out_nchw.shape[3]
        self._keep = (src0, src1, wt, out, bias, bias_img, film, ln_g, res, res2, out_nchw, res_f32, out_f32, out_planar)
        self.desc = d
        h = C.c_void_p()
        L.check(lib.dac_conv_create(C.byref(d), C.byref(h)))
        self.handle = h
        self._lib = lib
        self._keep += (rsrc0, rsrc1, rweight, stats_out, ln_stats, ln_colsum, kv_shift, ctx_acc)
        self.flops = 2.0 * B * OH * OW * pw.ngroups * len(pw.taps[0]) * (c0 + c1) * pw.cout \
            + 2.0 * B * OH * OW * (rc0 + rc1) * pw.cout * (rsrc0 is not None)
        # algorithmic DRAM bytes of one launch: every operand tensor once (bf16 unless noted)
        opix = B * OH * OW * pw.out_scale ** 2
        out_c = pw.cout // 2 if epi == L.EPI_GEGLU else pw.cout
        self.bytes = (2.0 * B * H * W * (c0 + c1) + 2.0 * wt.numel()
                      + (2.0 * opix * out_c if out is not None else 0.0)
                      + (4.0 * out_nchw.numel() if out_nchw is not None else 0.0)
                      + (4.0 * out_f32.numel() if out_f32 is not None else 0.0)
                      + (4.0 * res_f32.numel() if res_f32 is not None else 0.0)
                      + (2.0 * out_planar.numel() if out_planar is not None else 0.0)
                      + 2.0 * opix * out_c * ((res is not None) + (res2 is not None))
                      + 2.0 * B * H * W * (rc0 + rc1) * (rsrc0 is not None))

    def run(self):
        L.check(self._lib.dac_conv_launch(self.handle, L.stream_ptr()))

    def info(self):
        v = [C.c_int32() for _ in range(4)]
        L.check(self._lib.dac_conv_info(self.handle, *[C.byref(x) for x in v]))
        return dict(tiles=v[0].value, ctas=v[1].value, smem=v[2].value, stages=v[3].value)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.dac_conv_destroy(self.handle)
        except Exception:
            pass


def pack_conv_pair(weight):
    """3x3 stride-1 conv weight [64, cin, 3, 3] for the pixel-pair mode of the conv kernel (dac_conv_desc.pair): per ky
    one 192-row block [W(kx=2); W(kx=1); W(kx=0)], K-major [3][192][cin] bf16.  The even-pixel chunk reads rows [64,192)
    (centre of the even output | left neighbour of the odd one) and rows [0,64) (right neighbour of the odd output, from
    the next pair); the odd-pixel chunk rows [0,128) and rows [128,192)."""
    cout, cin, kh, kw = weight.shape
    assert (kh, kw) == (3, 3) and cin % 64 == 0 and cout in (16, 64), "64 output channels, or final_conv padded to 16"
    w = torch.stack([torch.cat([weight[:, :, ky, 2], weight[:, :, ky, 1], weight[:, :, ky, 0]], 0) for ky in range(3)])
    if cout != 64:
        return w.to(torch.bfloat16).contiguous()
    # ... followed by the two per-rank layouts of the CTA-pair build (conv_kernel.cuh, cta_group::2: CTA r supplies the weight
    # rows of output columns [64 r, 64 r + 64) of the centre windows and rows [32 r, 32 r + 32) of the side windows):
    #   rank 0: [W(1); W(2); W(0)[0:32]; W(2)[0:32]]      rank 1: [W(0); W(1); W(0)[32:64]; W(2)[32:64]]
    blocks = [w]
    for r in (0, 1):
        blocks.append(torch.stack([torch.cat([weight[:, :, ky, 1 - r], weight[:, :, ky, 2 - r],
                                              weight[32 * r:32 * r + 32, :, ky, 0], weight[32 * r:32 * r + 32, :, ky, 2]], 0)
                                   for ky in range(3)]))
    return torch.cat(blocks, 0).to(torch.bfloat16).contiguous()


def pair_eligible(pw, cout, c0, c1, W):
    """3x3 stride-1 convs with 64 output channels, 64-channel sources and an even width (see PairConvPlan)."""
    return (cout == 64 and c0 == 64 and c1 in (0, 64) and W % 2 == 0 and W >= 16 and not os.environ.get("DAC_NO_PAIR"))


class PairConvPlan(ConvPlan):
    """A 64-output-channel 3x3 conv in pixel-pair mode: every NHWC tensor [B, H, W, C] is handed to the kernel as its
    [B, H, W/2, 2C] view, so a GEMM row is a pair of adjacent pixels and the centre taps run as N = 128 MMAs (an N = 64
    MMA is capped at 66.6 % of the tensor peak by its shared-memory operand fetch, profiles/r01_mma_rate.txt)."""

    def __init__(self, src0, wpair, out, *, B, H, W, src1=None, act=L.ACT_NONE, film=None, film_off=0, res=None,
                 rsrc0=None, rsrc1=None, rweight=None, out_nchw=None, bias=None):
        L.require_cuda(src0, wpair, out if out is not None else out_nchw)
        assert W % 2 == 0 and src0.shape[-1] == 64 and (out is None or out.shape[-1] == 64)
        bn = 32 if out_nchw is not None else 128      # final_conv: 2 x 16 padded output channels, fp32 NCHW, cropped
        lib = L.load()
        d = L.ConvDesc()
        Wp = W // 2
        d.src0, d.c0, d.ld0 = src0.data_ptr(), 128, 128
        if src1 is not None:
            assert src1.shape[-1] == 64
            d.src1, d.c1, d.ld1 = src1.data_ptr(), 128, 128
        d.B, d.H, d.W, d.OH, d.OW = B, H, Wp, H, Wp
        d.stride, d.ngroups, d.ntaps, d.ndy, d.ncols = 1, 1, 9, 9, 1
        d.col_dx[0][0], d.col_dy0[0][0] = -1, -1
        for i in range(9):
            d.col_tap[0][i] = i
        d.out_scale = 1
        d.weight, d.cout, d.cout_pad, d.per_image_w = wpair.data_ptr(), bn, bn, 0
        d.block_n, d.tile_h, d.tile_w = bn, 16, 8
        assert wpair.shape[-2] == 3 * bn // 2
        d.epi, d.act, d.halo, d.pair = L.EPI_PLAIN, act, 1, (2 if wpair.shape[0] == 9 else 1)
        if film is not None:
            d.film, d.film_ld, d.film_off = film.data_ptr(), film.shape[-1], film_off
        if res is not None:
            assert res.shape[-1] == 64
            d.res, d.res_ld = res.data_ptr(), 128
        if out is not None:
            d.out, d.out_ld, d.out_coff = out.data_ptr(), 128, 0
        if out_nchw is not None:
            d.out_nchw = out_nchw.data_ptr()
            d.out_nchw_c, d.out_nchw_h, d.out_nchw_w = out_nchw.shape[1], out_nchw.shape[2], out_nchw.shape[3]
        if bias is not None:
            d.bias = bias.data_ptr()
        rc = 0
        if rsrc0 is not None:        # fused 1x1 skip conv over (rsrc0 | rsrc1), 64 channels each: rweight = pack_linear([64, rc])
            assert rsrc0.shape[-1] == 64 and rweight.w.shape[-2] == 64
            d.rsrc0, d.rc0, d.rld0 = rsrc0.data_ptr(), 128, 128
            rc = 64
            if rsrc1 is not None:
                assert rsrc1.shape[-1] == 64
                d.rsrc1, d.rc1, d.rld1 = rsrc1.data_ptr(), 128, 128
                rc = 128
            d.rweight = rweight.w.data_ptr()
        self._keep = (src0, src1, wpair, out, film, res, rsrc0, rsrc1, rweight, out_nchw, bias)
        self.desc = d
        h = C.c_void_p()
        L.check(lib.dac_conv_create(C.byref(d), C.byref(h)))
        self.handle = h
        self._lib = lib
        cout = out_nchw.shape[1] if out_nchw is not None else 64
        self.flops = 2.0 * B * H * W * 9 * (64 + (64 if src1 is not None else 0)) * cout + 2.0 * B * H * W * rc * 64
        self.bytes = (2.0 * B * H * W * 64 * (1 + (src1 is not None)) + 2.0 * wpair.numel()
                      + (4.0 * out_nchw.numel() if out_nchw is not None else 2.0 * B * H * W * 64)
                      + 2.0 * B * H * W * 64 * (res is not None) + 2.0 * B * H * W * rc)


def ctx_slots(B, h, w, tensor_core_kv):
    """Partial context records per (image, head) the k|v kernels store (dac_linattn_ctx_slots): one per CTA that touches
    the image for dac_linattn_kv, two (one per epilogue group) for the KVCTX convolution epilogue."""
    if tensor_core_kv:
        tpi, groups = (h * w) // 128, 1
    else:
        th, tw = choose_tile(h, w, 1)
        tpi, groups = (-(-h // th)) * (-(-w // tw)), 2
    return int(L.load().dac_linattn_ctx_slots(B, tpi, groups))


def pack_kv_grouped(wkv):
    """[256, C] rows (k heads 0-3 | v heads 0-3, 32 each) -> rows packed per head pair g: k_2g k_2g+1 v_2g v_2g+1
    (the N = 128 accumulator of epilogue group g then holds exactly its heads); bf16 contiguous."""
    k, v = wkv[:128].reshape(2, 64, -1), wkv[128:].reshape(2, 64, -1)
    return torch.cat([k[0], v[0], k[1], v[1]], 0).to(torch.bfloat16).contiguous()


class KvPlan:
    """LinearAttention key/value side on tcgen05: k|v GEMM, exp, context GEMM accumulated in tensor memory."""

    def __init__(self, xn, wkv_grouped, kv_shift, ctx_acc, B, hw, Cn, ln_stats=None, ln_colsum=None, prenorm_eps=None):
        """prenorm_eps: `xn` is the RAW tensor and the kernel applies the gain-free channel LayerNorm itself (C = 64); then
        `wkv_grouped` holds only the ROW-CENTRED key rows ([128, 64] bf16, head order; centre_rows) and `ctx_acc` is [B, 4, slots, KV_G_REC] for
        linattn_fold_g."""
        L.require_cuda(xn, wkv_grouped, kv_shift, ctx_acc, ln_stats, ln_colsum)
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.dac_linattn_kv_create(xn.data_ptr(), wkv_grouped.data_ptr(), kv_shift.data_ptr(),
                                          ctx_acc.data_ptr(), ctx_acc.shape[-2], L.ptr(ln_stats), L.ptr(ln_colsum),
                                          B, hw, Cn, int(prenorm_eps is not None), float(prenorm_eps or 0.0), C.byref(h)))
        self.handle, self._lib = h, lib
        self._keep = (xn, wkv_grouped, kv_shift, ctx_acc, ln_stats, ln_colsum)
        self.flops = 2.0 * B * hw * Cn * 256      # algorithmic (k and v rows), whichever way the kernel gets there
        self.bytes = 2.0 * B * hw * Cn + 4.0 * ctx_acc.numel()

    def run(self):
        L.check(self._lib.dac_linattn_kv_launch(self.handle, L.stream_ptr()))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.dac_linattn_kv_destroy(self.handle)
        except Exception:
            pass


class QoutPlan:
    """LinearAttention query side (to_q softmax -> W_eff q -> LayerNorm -> + x) as one chained-GEMM launch."""

    def __init__(self, xn, wq, weff, res, out, bias, ln_g, ln_eps, B, hw, Cn, ln_stats=None, ln_colsum=None,
                 prenorm_eps=None, q_shift=None):
        """prenorm_eps: `xn` (== `res`) is the RAW tensor, `wq` the ROW-CENTRED weight (centre_rows) and the kernel folds the
        gain-free channel LayerNorm into its first GEMM (C = 64); q_shift: [128] fp32 bounds c_d log2(e) for a softmax shift
        without the row maximum."""
        L.require_cuda(xn, wq, weff, res, out, ln_g, ln_stats, ln_colsum, q_shift)
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.dac_linattn_qout_create(xn.data_ptr(), wq.data_ptr(), weff.data_ptr(), weff.shape[-2],
                                            res.data_ptr(), out.data_ptr(), bias.data_ptr() if bias is not None else None,
                                            ln_g.data_ptr(), ln_eps, L.ptr(ln_stats), L.ptr(ln_colsum), B, hw, Cn,
                                            int(prenorm_eps is not None), float(prenorm_eps or 0.0), L.ptr(q_shift),
                                            C.byref(h)))
        self.handle, self._lib = h, lib
        self._keep = (xn, wq, weff, res, out, bias, ln_g, ln_stats, ln_colsum, q_shift)
        self.flops = 2.0 * B * hw * 128 * Cn * 2
        self.bytes = 2.0 * B * hw * Cn * (2 if xn.data_ptr() == res.data_ptr() else 3) + 2.0 * weff.numel()

    def run(self):
        L.check(self._lib.dac_linattn_qout_launch(self.handle, L.stream_ptr()))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.dac_linattn_qout_destroy(self.handle)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ thin wrappers
def _coef(vals):
    arr = (C.c_float * 8)()
    for i, v in enumerate(vals):
        arr[i] = float(v)
    return arr


def _same_fp32(what, ref, *tensors):
    """Raw-pointer kernels: every operand must be fp32, contiguous, on one device and exactly as large as the state."""
    for name, t in tensors:
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != ref.numel() or t.device != ref.device:
            raise L.DacError(f"{what}: `{name}` must be a contiguous fp32 tensor of {ref.numel()} elements on {ref.device} "
                             f"(got {t.dtype}, {tuple(t.shape)}, {t.device}, contiguous={t.is_contiguous()})")


def sde_step(mode, x, mu, net, eps, out, coef):
    """out = one reverse step from x (out may alias x: each element is read before it is written)."""
    L.require_cuda(x, mu, net, eps, out)
    _same_fp32("sde_step", x, ("x", x), ("mu", mu), ("net", net), ("eps", eps), ("out", out))
    if mode != 2 and eps is None:
        raise L.DacError("sde_step: the sde / posterior updates need a noise tensor")
    L.check(L.load().dac_sde_step(mode, L.ptr(x), L.ptr(mu), L.ptr(net), L.ptr(eps), L.ptr(out), x.numel(),
                                  _coef(coef), L.stream_ptr()))


def noise_state(x, eps, out, max_sigma):
    L.require_cuda(x, eps, out)
    _same_fp32("noise_state", x, ("x", x), ("eps", eps), ("out", out))
    L.check(L.load().dac_noise_state(L.ptr(x), L.ptr(eps), L.ptr(out), x.numel(), float(max_sigma), L.stream_ptr()))


def loop_tick(state, t_table, coef_table, t_dev, coef_dev):
    """First launch of a device-driven sampling step (dac_loop_tick): publishes the step's time and coefficients."""
    L.check(L.load().dac_loop_tick(L.ptr(state), L.ptr(t_table), L.ptr(coef_table), L.ptr(t_dev), L.ptr(coef_dev),
                                   L.stream_ptr()))


def sde_step_dev(mode, x, mu, net, out, coef_dev, state):
    """Last launch of a device-driven sampling step (dac_sde_step_dev)."""
    _same_fp32("sde_step_dev", x, ("x", x), ("mu", mu), ("net", net), ("out", out))
    L.check(L.load().dac_sde_step_dev(mode, L.ptr(x), L.ptr(mu), L.ptr(net), L.ptr(out), x.numel(), L.ptr(coef_dev),
                                      L.ptr(state), L.stream_ptr()))


def stem_input(xt, cond, out, H, W, pair=False):
    """out: [B, Hp, Wp, 64], or [B, Hp, Wp/2, 64] with pair=True (one packed row per pair of adjacent pixels)."""
    B, Hp, Wp = out.shape[0], out.shape[1], out.shape[2] * (2 if pair else 1)
    L.check(L.load().dac_unet_stem_input(L.ptr(xt), L.ptr(cond), L.ptr(out), B, H, W, Hp, Wp, int(pair), L.stream_ptr()))


def layernorm_rows(x, out, rows, c, w=None, b=None, eps=1e-5):
    L.check(L.load().dac_layernorm_rows(L.ptr(x), x.shape[-1], L.ptr(out), out.shape[-1], rows, c, L.ptr(w), L.ptr(b),
                                        float(eps), L.stream_ptr()))


def layernorm_rows_res(x, res, out, rows, c, w=None, b=None, eps=1e-5):
    """out = LayerNorm(x) * w + b + res  (bf16 rows; res: bf16 rows of the same width)."""
    L.check(L.load().dac_layernorm_rows_res(L.ptr(x), x.shape[-1], L.ptr(res), res.shape[-1], L.ptr(out), out.shape[-1], rows,
                                            c, L.ptr(w), L.ptr(b), float(eps), L.stream_ptr()))


def layernorm_rows_f32(x, out, rows, c, w=None, b=None, eps=1e-5):
    L.check(L.load().dac_layernorm_rows_f32(L.ptr(x), x.shape[-1], L.ptr(out), out.shape[-1], rows, c, L.ptr(w),
                                            L.ptr(b), float(eps), L.stream_ptr()))


def groupnorm_nhwc(x, out, B, hw, c, w, b, stats, groups=32, eps=1e-6):
    L.check(L.load().dac_groupnorm_nhwc(L.ptr(x), L.ptr(out), B, hw, c, groups, L.ptr(w), L.ptr(b), float(eps),
                                        L.ptr(stats), L.stream_ptr()))


def prenorm_groupnorm_nhwc(x, normed, out, B, hw, c, pre_g, w, b, stats, groups=32, pre_eps=1e-5, eps=1e-6):
    """normed = LayerNorm_c(x) * pre_g, out = GroupNorm(normed): the PreNorm + Normalize pair in front of a SpatialTransformer."""
    L.check(L.load().dac_prenorm_groupnorm_nhwc(L.ptr(x), L.ptr(normed), L.ptr(out), B, hw, c, groups, L.ptr(pre_g),
                                                float(pre_eps), L.ptr(w), L.ptr(b), float(eps), L.ptr(stats), L.stream_ptr()))


def prompt_embed(ew, text_ctx, B, prompt_emb):
    L.check(L.load().dac_prompt_embed(C.byref(ew), L.ptr(text_ctx), B, L.ptr(prompt_emb), L.stream_ptr()))


def time_film(ew, time, prompt_emb, B, temb, film):
    L.check(L.load().dac_time_film(C.byref(ew), L.ptr(time), L.ptr(prompt_emb), B, L.ptr(temb), L.ptr(film),
                                   L.stream_ptr()))


def two_linear(x, w1, w2, b2, y):
    B, kin = x.shape
    L.check(L.load().dac_two_linear(L.ptr(x), B, kin, L.ptr(w1), w1.shape[0], L.ptr(w2), L.ptr(b2), w2.shape[0],
                                    L.ptr(y), L.stream_ptr()))


def linattn_context(kv, B, hw, nchunks, partial):
    L.check(L.load().dac_linattn_context(L.ptr(kv), B, hw, nchunks, L.ptr(partial), L.stream_ptr()))


def linattn_fold(partial, B, hw, nchunks, w_out, C_, c_pad, weff):
    L.check(L.load().dac_linattn_fold(L.ptr(partial), B, hw, nchunks, L.ptr(w_out), C_, c_pad, L.ptr(weff),
                                      L.stream_ptr()))


def linattn_fold_g(partial, B, hw, nslots, m_fold, C_, c_pad, weff):
    """Fold for KvPlan(prenorm_eps=...): records {G[32][64], S[32]} -> weff (dac_linattn_fold_g)."""
    L.check(L.load().dac_linattn_fold_g(L.ptr(partial), B, hw, nslots, L.ptr(m_fold), C_, c_pad, L.ptr(weff),
                                        L.stream_ptr()))


def centre_rows(w):
    """W - rowmean(W): W_c x = W (x - mean(x)), so a GEMM on the RAW tensor followed by the pixel's rstd equals W LN(x)."""
    return w - w.mean(dim=1, keepdim=True)


def kv_fold_matrix(w_out, w_v):
    """M_h[c'][c] = sum_e W_out[c'][h*32+e] W_vc[h*32+e][c] (fp32 [4, C, Cin]) from to_out's [C, 128] weight and the
    gain-folded value rows [128, Cin] of to_qkv - row-centred and rounded to bf16 as the GEMM operand would have been; the
    rows of M_h are centred again in fp32 so that the per-pixel mean drops out of G' M_h^T exactly."""
    wv = centre_rows(w_v.float()).to(torch.bfloat16).float().reshape(4, 32, -1)
    wo = w_out.float().reshape(w_out.shape[0], 4, 32).permute(1, 0, 2)
    m = torch.bmm(wo, wv)
    return (m - m.mean(dim=2, keepdim=True)).contiguous()


KV_G_REC = 32 * 64 + 32      # floats per {G, S} record of the in-kernel-PreNorm k kernel


def attention(qkv, out, B, n, heads, d):
    L.check(L.load().dac_attention(L.ptr(qkv), L.ptr(out), B, n, heads, d, L.stream_ptr()))


def vit_patchify(image, out, B, S, p):
    L.check(L.load().dac_vit_patchify(L.ptr(image), L.ptr(out), B, S, p, out.shape[-1], L.stream_ptr()))


def vit_embed(patch_emb, cls, pos, ln_w, ln_b, out, B, Ltok, w, eps=1e-5):
    L.check(L.load().dac_vit_embed(L.ptr(patch_emb), L.ptr(cls), L.ptr(pos), L.ptr(ln_w), L.ptr(ln_b), L.ptr(out),
                                   B, Ltok, w, float(eps), L.stream_ptr()))


def vit_pool(x, B, Ltok, w, ln_w, ln_b, proj, out, eps=1e-5):
    L.check(L.load().dac_vit_pool(L.ptr(x), B, Ltok, w, L.ptr(ln_w), L.ptr(ln_b), float(eps), L.ptr(proj),
                                  proj.shape[1], L.ptr(out), L.stream_ptr()))


def attention_causal(qkv, out, B, n, heads, d):
    L.check(L.load().dac_attention_causal(L.ptr(qkv), L.ptr(out), B, n, heads, d, L.stream_ptr()))


def text_embed(text, token_embedding, pos, out, eot, B, Ltok, w):
    L.check(L.load().dac_text_embed(L.ptr(text), L.ptr(token_embedding), L.ptr(pos), L.ptr(out), L.ptr(eot), B, Ltok, w,
                                    token_embedding.shape[0], L.stream_ptr()))


def text_pool(x, eot, B, Ltok, w, ln_w, ln_b, proj, out, eps=1e-5):
    L.check(L.load().dac_text_pool(L.ptr(x), L.ptr(eot), B, Ltok, w, L.ptr(ln_w), L.ptr(ln_b), float(eps), L.ptr(proj),
                                   proj.shape[1], L.ptr(out), L.stream_ptr()))


def degradation_argmax(degra, text, logits, argmax):
    B, e = degra.shape
    L.check(L.load().dac_degradation_argmax(L.ptr(degra), L.ptr(text), B, e, text.shape[0], L.ptr(logits),
                                            L.ptr(argmax), L.stream_ptr()))
