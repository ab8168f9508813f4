"""CPU emulation of the C ABI for HOST-LOGIC tests only (tests/ infrastructure).

The product has no CPU path.  This object is installed with ``_lib.set_backend`` by the
``fake_backend`` fixture so that the Python side (layout/pitch bookkeeping, autograd
formulas, module wiring, engine, optimizer arena, data-parallel plumbing) can be exercised
on a box without a GPU.  Each entry point is emulated with stock torch CPU ops on the very
tensors (views) the real backend would turn into pointers, with the same in-place /
accumulate semantics as the kernels documented in ``include/tss_b200.h``.
"""
import math

import torch
import torch.nn.functional as F
from torch.nn import grad as nngrad

RELU = 1


def _epilogue(y, scale, shift, res, flags):
    if shift is not None:
        s = scale.view(1, -1, 1, 1) if scale is not None else 1.0
        y = y * s + shift.view(1, -1, 1, 1)
    if res is not None:
        y = y + res.float()
    if flags & RELU:
        y = y.clamp_min(0)
    return y


def _stats(stats, y):
    if stats is not None:
        C = y.shape[1]
        stats[:C] += y.sum(dim=(0, 2, 3))
        stats[C:2 * C] += (y * y).sum(dim=(0, 2, 3))


class FakeBackend:
    def __init__(self):
        self.launches = 0

    def call(self, name, k):
        self.launches += 1
        return getattr(self, name)(**k)

    # ---------------------------------------------------------------- depthwise
    def tss_dwconv3x3_fwd(self, x, w, y, N, Hi, Wi, C, stride, dilation, scale, shift, flags, stats, dtype):
        assert x.shape == (N, C, Hi, Wi)
        raw = F.conv2d(x.float(), w.view(C, 1, 3, 3), None, stride, dilation, dilation, C)
        _stats(stats, raw)
        y.copy_(_epilogue(raw, scale, shift, None, flags))
        return 0

    # The kernels that apply the producer's BatchNorm while reading never round the activated value to the storage
    # type (they are more accurate than apply-then-read).  round_in_act = True emulates that rounding, which makes the
    # hand-over paths bit-identical to the default path: a plumbing check without bf16 noise.
    round_in_act = False

    def _in_act(self, x, in_scale, in_shift, in_flags):
        z = x.float() * in_scale.view(1, -1, 1, 1) + in_shift.view(1, -1, 1, 1)
        z = z.clamp_min(0) if in_flags & RELU else z
        return z.to(x.dtype).float() if self.round_in_act else z

    def tss_dwconv3x3_fwd_bnin(self, x, in_scale, in_shift, in_flags, w, y, N, Hi, Wi, C, stride, stats, dtype):
        raw = F.conv2d(self._in_act(x, in_scale, in_shift, in_flags), w.view(C, 1, 3, 3), None, stride, 1, 1, C)
        _stats(stats, raw)
        y.copy_(raw)
        return 0

    def tss_dwconv3x3_wgrad_bnin(self, x, in_scale, in_shift, in_flags, dy, dw, N, Hi, Wi, C, stride, dtype):
        dw += nngrad.conv2d_weight(self._in_act(x, in_scale, in_shift, in_flags), (C, 1, 3, 3), dy.float(), stride, 1, 1, C).view_as(dw)
        return 0

    def tss_dwconv3x3_dgrad(self, dy, w, dx, N, Hi, Wi, C, stride, dilation, dtype):
        dx.copy_(nngrad.conv2d_input((N, C, Hi, Wi), w.view(C, 1, 3, 3), dy.float(), stride, dilation, dilation, C))
        return 0

    def tss_dwconv3x3_wgrad(self, x, dy, dw, N, Hi, Wi, C, stride, dilation, dtype):
        dw += nngrad.conv2d_weight(x.float(), (C, 1, 3, 3), dy.float(), stride, dilation, dilation, C).view_as(dw)
        return 0

    # ---------------------------------------------------------------- pointwise
    def tss_pwconv_fwd(self, x, w, wp, y, M, K, Nc, ldx, ldy, scale, shift, res, ldr, flags, stats, impl, dtype):
        assert x.shape[1] == K and y.shape[1] == Nc and x.shape[0] * x.shape[2] * x.shape[3] == M
        assert x.stride(3) == ldx or x.shape[3] == 1
        w_eff = wp.float() if (impl == 1 and wp is not None) else w.float()      # impl 1 multiplies with the bf16 pack
        raw = F.conv2d(x.float(), w_eff.view(Nc, K, 1, 1))
        _stats(stats, raw)
        y.copy_(_epilogue(raw, scale, shift, res, flags))
        return 0

    def tss_pwconv_fwd_bnin(self, x, ldx, in_scale, in_shift, in_flags, z, ldz, wp, y, ldy, M, K, Nc, stats):
        act = self._in_act(x, in_scale, in_shift, in_flags)
        if z is not None:
            z.copy_(act)
            act = z.float()                       # the GEMM multiplies the rounded operand
        raw = F.conv2d(act, wp.float().view(Nc, K, 1, 1))
        _stats(stats, raw)
        y.copy_(raw)
        return 0

    def tss_pwconv_dgrad(self, dy, w, wpT, dx, M, K, Nc, lddy, lddx, impl, dtype):
        wt = wpT.float() if (impl == 1 and wpT is not None) else w.view(Nc, K).t()      # impl 1 multiplies with the bf16 pack
        dx.copy_(F.conv2d(dy.float(), wt.reshape(K, Nc, 1, 1)))
        return 0

    def _bnred(self, dz, g, yp, mean, rstd, gamma, beta, flags, sums):
        v = lambda t: t.detach().view(1, -1, 1, 1)
        C = dz.shape[1]
        if flags & RELU:
            sc = v(gamma) * v(rstd)
            dz = dz * (torch.addcmul(v(beta) - v(mean) * sc, yp.float(), sc) > 0)
        sums[:C] += dz.sum(dim=(0, 2, 3))
        sums[C:] += (dz * (yp.float() - v(mean)) * v(rstd)).sum(dim=(0, 2, 3))
        g.copy_(dz)
        return 0

    def tss_pwconv_dgrad_bnred(self, dy, wpT, g, M, K, Nc, lddy, ldg, yp, ldyp, mean, rstd, gamma, beta, flags, sums):
        dz = F.conv2d(dy.float(), wpT.float().reshape(K, Nc, 1, 1))
        return self._bnred(dz, g, yp, mean, rstd, gamma, beta, flags, sums)

    def tss_pwconv_bwd_fused(self, dz, y, lddz, ldy, mean, rstd, gamma, beta, sums, flags, count, dy, lddy, dgamma,
                             dbeta, wpT, dx, M, K, Nc, lddx, yp, ldyp, pmean, prstd, pgamma, pbeta, pflags, psums):
        self.tss_bn_bwd_apply(dz, None, y, mean, rstd, gamma, beta, sums, dy, None, dgamma, dbeta, M, count, Nc, lddz, 0,
                              ldy, lddy, 0, flags, 0)
        out = F.conv2d(dy.float(), wpT.float().reshape(K, Nc, 1, 1))
        if yp is None:
            dx.copy_(out)
            return 0
        return self._bnred(out, dx, yp, pmean, prstd, pgamma, pbeta, pflags, psums)

    def tss_dwconv3x3_dgrad_bnred(self, dy, w, g, N, H, W, C, yp, mean, rstd, gamma, beta, flags, sums, dtype):
        dz = nngrad.conv2d_input((N, C, H, W), w.detach().view(C, 1, 3, 3), dy.float(), 1, 1, 1, C)
        return self._bnred(dz, g, yp, mean, rstd, gamma, beta, flags, sums)

    def tss_dwconv3x3_bwd_fused(self, dz, y, w, mean, rstd, gamma, beta, sums, flags, count, dy, dgamma, dbeta, g, N, H, W, C,
                                yp, pmean, prstd, pgamma, pbeta, pflags, psums, dtype):
        out = dy if dy is not None else torch.empty_like(dz)
        self.tss_bn_bwd_apply(dz, None, y, mean, rstd, gamma, beta, sums, out, None, dgamma, dbeta, N * H * W, count, C, C, 0, C,
                              C, 0, flags, dtype)
        d = nngrad.conv2d_input((N, C, H, W), w.detach().view(C, 1, 3, 3), out.float(), 1, 1, 1, C)
        if yp is None:
            g.copy_(d)
            return 0
        return self._bnred(d, g, yp, pmean, prstd, pgamma, pbeta, pflags, psums)

    def tss_dwconv3x3_dgrad_s2_bnred(self, dy, w, g, N, Hi, Wi, C, yp, mean, rstd, gamma, beta, flags, sums, dtype):
        dz = nngrad.conv2d_input((N, C, Hi, Wi), w.detach().view(C, 1, 3, 3), dy.float(), 2, 1, 1, C)
        return self._bnred(dz, g, yp, mean, rstd, gamma, beta, flags, sums)

    def tss_pwconv_wgrad(self, x, dy, dw, db, M, K, Nc, ldx, lddy, impl, dtype):
        g = dy.float().permute(1, 0, 2, 3).reshape(Nc, -1)
        a = x.float().permute(1, 0, 2, 3).reshape(K, -1)
        dw += (g @ a.t()).view_as(dw)
        if db is not None:
            db += g.sum(1)
        return 0

    def tss_pack_weights_bf16(self, w, wp, wpT, Nc, K):
        if wp is not None:
            wp.copy_(w.view(Nc, K))
        if wpT is not None:
            wpT.copy_(w.view(Nc, K).t())
        return 0

    def tss_class_scores_pack(self, w, bias, wp, wpT, bias_pad, Nc, K, Np, Npt):
        wp.zero_(); wpT.zero_(); bias_pad.zero_()
        wp[:Nc].copy_(w.view(Nc, K))
        wpT[:, :Nc].copy_(w.view(Nc, K).t())
        if bias is not None:
            bias_pad[:Nc].copy_(bias)
        return 0

    def tss_dwpw_fwd(self, x, w_dw, scale1, shift1, flags1, wp, y, N, H, W, C, Nc, ldy, scale2, shift2, res, ldr, flags2):
        t = F.conv2d(x.float(), w_dw.detach().view(C, 1, 3, 3), None, 1, 1, 1, C)
        t = _epilogue(t, scale1, shift1, None, flags1).to(x.dtype).float()
        raw = F.conv2d(t, wp.float().view(Nc, C, 1, 1))
        y.copy_(_epilogue(raw, scale2, shift2, res, flags2))
        return 0

    def tss_pack_weights_multi(self, arena, table, n_entries, max_elems):
        import ctypes
        for off, Nc, K, wp, wpT in table.tolist():
            w = arena[off:off + Nc * K].view(Nc, K)
            for addr, src in ((wp, w), (wpT, w.t())):
                if addr:
                    dst = torch.frombuffer((ctypes.c_uint16 * (Nc * K)).from_address(int(addr)), dtype=torch.bfloat16)
                    dst.copy_(src.reshape(-1))
        return 0

    # ---------------------------------------------------------------- dense 3x3 as a patch GEMM
    def tss_im2col3x3(self, x, col, N, H, W, C, dtype):
        xp = F.pad(x, (1, 1, 1, 1))
        taps = [xp[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)]
        col.copy_(torch.cat(taps, dim=1))
        return 0

    def tss_col2im3x3(self, dcol, dx, N, H, W, C, dtype):
        acc = torch.zeros(N, C, H + 2, W + 2)
        for tap in range(9):
            ky, kx = divmod(tap, 3)
            acc[:, :, ky:ky + H, kx:kx + W] += dcol[:, tap * C:(tap + 1) * C].float()
        dx.copy_(acc[:, :, 1:H + 1, 1:W + 1])
        return 0

    def tss_permute_weights3x3(self, src, dst, Cout, Cin, backward):
        if backward:
            dst += src.detach().view(Cout, 9, Cin).permute(0, 2, 1).reshape(Cout, Cin, 3, 3)
        else:
            dst.copy_(src.detach().reshape(Cout, Cin, 9).permute(0, 2, 1).reshape(Cout, 9 * Cin, 1, 1))
        return 0

    # ---------------------------------------------------------------- dropout
    def _dropout(self, x, y, n, p, seed, offset):
        from tests.philox_ref import keep_mask
        keep, scale = keep_mask(int(seed), int(offset), n, p)
        flat = x.permute(0, 2, 3, 1).reshape(-1).float()              # NHWC element order
        out = torch.where(torch.from_numpy(keep), flat * float(scale), torch.zeros_like(flat))
        y.copy_(out.view(x.shape[0], x.shape[2], x.shape[3], x.shape[1]).permute(0, 3, 1, 2))
        return 0

    def tss_dropout_fwd(self, x, y, n, p, rng, used, dtype):
        assert int(rng[2]) == 0
        used[0] = rng[1]
        rng[1] += 1
        return self._dropout(x, y, n, p, rng[0], used[0])

    def tss_dropout_bwd(self, dy, dx, n, p, rng, used, dtype):
        return self._dropout(dy, dx, n, p, rng[0], used[0])

    # ---------------------------------------------------------------- pyramid pooling, grouped
    @staticmethod
    def _at(addr, n, dtype=torch.float32):
        """A tensor over ``n`` elements at a host address taken from the kernels' address table."""
        import ctypes
        if addr == 0:
            return None
        ctype = ctypes.c_float if dtype == torch.float32 else ctypes.c_int64
        return torch.frombuffer((ctype * n).from_address(int(addr)), dtype=dtype)

    @staticmethod
    def _branch_rows(N, bins):
        off = 0
        for b in bins.values:
            yield b, off, off + N * b * b
            off += N * b * b

    def tss_ppm_branches_fwd(self, pool, table, y, z, mean, rstd, N, C, Cb, bins, nbins, momentum, eps, dtype):
        for i, (b, lo, hi) in enumerate(self._branch_rows(N, bins)):
            t = table[i].tolist()
            w, gamma, beta = self._at(t[0], Cb * C).view(Cb, C), self._at(t[1], Cb), self._at(t[2], Cb)
            raw = pool[lo:hi].float() @ w.t()
            M = hi - lo
            assert M > 1
            mu = raw.double().mean(0)
            var = (raw.double() ** 2).mean(0) - mu * mu
            rs = (1.0 / torch.sqrt(var.clamp_min(0) + eps)).float()
            mean[i].copy_(mu.float())
            rstd[i].copy_(rs)
            y[lo:hi].copy_(raw)
            sc = gamma * rs
            z[lo:hi].copy_(torch.addcmul(beta - mu.float() * sc, y[lo:hi].float(), sc).clamp_min(0))
            rm, rv, nbt = self._at(t[3], Cb), self._at(t[4], Cb), self._at(t[5], 1, torch.int64)
            if rm is not None:
                rm.mul_(1 - momentum).add_(momentum * mu.float())
                rv.mul_(1 - momentum).add_(momentum * (var.clamp_min(0) * M / (M - 1)).float())
            if nbt is not None:
                nbt += 1
        return 0

    def tss_ppm_branches_eval(self, pool, table, z, N, C, Cb, bins, nbins, dtype):
        for i, (b, lo, hi) in enumerate(self._branch_rows(N, bins)):
            t = table[i].tolist()
            w, sc, sh = self._at(t[0], Cb * C).view(Cb, C), self._at(t[1], Cb), self._at(t[2], Cb)
            z[lo:hi].copy_(torch.addcmul(sh, pool[lo:hi].float() @ w.t(), sc).clamp_min(0))
        return 0

    def tss_ppm_concat_fwd(self, x, z, cat, N, H, W, C, Cb, bins, nbins, dtype):
        cat[:, :C].copy_(x)
        for i, (b, lo, hi) in enumerate(self._branch_rows(N, bins)):
            zi = z[lo:hi].view(N, b, b, Cb).permute(0, 3, 1, 2).float()
            cat[:, C + i * Cb:C + (i + 1) * Cb].copy_(F.interpolate(zi, size=(H, W), mode='bilinear', align_corners=True))
        return 0

    def tss_ppm_concat_bwd(self, dcat, dz, N, H, W, C, Cb, lddcat, bins, nbins, dtype):
        for i, (b, lo, hi) in enumerate(self._branch_rows(N, bins)):
            with torch.enable_grad():
                zi = torch.zeros(N, Cb, b, b, requires_grad=True)
                up = F.interpolate(zi, size=(H, W), mode='bilinear', align_corners=True)
                (g,) = torch.autograd.grad(up, zi, dcat[:, C + i * Cb:C + (i + 1) * Cb].float())
            dz[lo:hi].copy_(g.permute(0, 2, 3, 1).reshape(hi - lo, Cb))
        return 0

    def tss_ppm_branches_bwd(self, dz, y, pool, table, mean, rstd, dy, dpool, N, C, Cb, bins, nbins, dtype):
        for i, (b, lo, hi) in enumerate(self._branch_rows(N, bins)):
            t = table[i].tolist()
            w, gamma, beta = self._at(t[0], Cb * C).view(Cb, C), self._at(t[1], Cb), self._at(t[2], Cb)
            M = hi - lo
            sc = gamma * rstd[i]
            yy = y[lo:hi].float()
            g = dz[lo:hi].float() * (torch.addcmul(beta - mean[i] * sc, yy, sc) > 0)
            xh = (yy - mean[i]) * rstd[i]
            s1, s2 = g.sum(0), (g * xh).sum(0)
            dy[lo:hi].copy_(sc * (g - s1 / M - xh * s2 / M))
            dw, dgamma, dbeta = self._at(t[6], Cb * C), self._at(t[7], Cb), self._at(t[8], Cb)
            if dbeta is not None:
                dbeta += s1
            if dgamma is not None:
                dgamma += s2
            d = dy[lo:hi].float()
            dpool[lo:hi].copy_(d @ w)
            if dw is not None:
                dw += (d.t() @ pool[lo:hi].float()).reshape(-1)
        return 0

    # ---------------------------------------------------------------- input pipeline
    def tss_augment_batch(self, images, labels, geom, lut, norm, out_image, out_label, N, H, W, ch, cw):
        import numpy as np
        from oracle import augment as A            # tests/ infrastructure may use the oracle
        mean, inv = np.array(norm.values[:3], dtype=np.float32), np.array(norm.values[3:], dtype=np.float32)
        for n in range(N):
            nh, nw, cy, cx, flip = (int(v) for v in geom[n])
            img = A.resize_linear_u8(images[n].numpy(), nh, nw)[cy:cy + ch, cx:cx + cw]
            if flip:
                img = img[:, ::-1]
            x = img.astype(np.float32)
            x -= mean
            x *= inv
            out_image[n].copy_(torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))))
            if labels is not None:
                lab = A.resize_nearest(labels[n].numpy(), nh, nw)[cy:cy + ch, cx:cx + cw]
                if flip:
                    lab = lab[:, ::-1]
                lab = torch.from_numpy(np.ascontiguousarray(lab)).long()
                out_label[n].copy_(lut[lab] if lut is not None else lab)
        return 0

    # ---------------------------------------------------------------- stem
    def tss_stem3x3s2_fwd(self, x, w, y, N, H, W, Cout, scale, shift, flags, stats, dtype):
        raw = F.conv2d(x, w, None, 2, 1)
        _stats(stats, raw)
        y.copy_(_epilogue(raw, scale, shift, None, flags))
        return 0

    def tss_stem3x3s2_fwd_tc(self, x, w, y, N, H, W, Cout, scale, shift, flags, stats):
        raw = F.conv2d(x.bfloat16().float(), w.detach().bfloat16().float(), None, 2, 1)      # operands rounded to bf16
        _stats(stats, raw)
        y.copy_(_epilogue(raw, scale, shift, None, flags))
        return 0

    def tss_stem3x3s2_wgrad_tc(self, x, dy, dw, N, H, W, Cout):
        dw += nngrad.conv2d_weight(x.bfloat16().float(), dw.shape, dy.float(), 2, 1)          # image rounded to bf16
        return 0

    def tss_stem3x3s2_patches(self, x, patches, N, H, W):
        cols = F.unfold(x, 3, padding=1, stride=2).transpose(1, 2).reshape(-1, 27)          # (ci, ky, kx) order
        patches.zero_()
        patches[:, :27].copy_(cols)
        return 0

    def tss_stem3x3s2_wgrad_from_patches(self, patches, dy, dw32, dw, M, Cout):
        g = dy.permute(0, 2, 3, 1).reshape(M, Cout).float()
        dw32.copy_(g.t() @ patches.float())
        dw += dw32[:, :27].reshape(dw.shape)
        return 0

    def tss_stem3x3s2_wgrad_patches(self, x, dy, patches, dw32, dw, N, H, W, Cout):
        self.tss_stem3x3s2_patches(x, patches, N, H, W)
        return self.tss_stem3x3s2_wgrad_from_patches(patches, dy, dw32, dw, patches.shape[0], Cout)

    def tss_stem3x3s2_wgrad_tc_bn(self, x, dz, y, mean, rstd, gamma, beta, sums, flags, count, dw, dgamma, dbeta, N, H, W, Cout):
        dy = torch.empty_like(dz)
        Ho, Wo = dz.shape[2], dz.shape[3]
        self.tss_bn_bwd_apply(dz, None, y, mean, rstd, gamma, beta, sums, dy, None, dgamma, dbeta, N * Ho * Wo, count, Cout, Cout, 0,
                              Cout, Cout, 0, flags, 1)
        return self.tss_stem3x3s2_wgrad_tc(x, dy, dw, N, H, W, Cout)

    def tss_stem3x3s2_wgrad(self, x, dy, dw, N, H, W, Cout, dtype):
        dw += nngrad.conv2d_weight(x, dw.shape, dy.float(), 2, 1)
        return 0

    # ---------------------------------------------------------------- batch norm
    def tss_bn_finalize(self, stats, count, gamma, beta, running_mean, running_var, num_batches_tracked,
                        momentum, eps, scale, shift, mean, rstd, C, clear_n=0):
        if count <= 1:
            raise RuntimeError('tss_bn_finalize failed (1): Expected more than 1 value per channel when training')
        m = stats[:C].double() / count
        var = (stats[C:2 * C].double() / count - m * m).clamp_min(0)
        r = 1.0 / torch.sqrt(var + eps)
        mean.copy_(m)
        rstd.copy_(r)
        scale.copy_(gamma.detach() * rstd)
        shift.copy_(beta.detach() - mean * scale)
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * m.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * count / (count - 1)).float())
        if num_batches_tracked is not None:
            num_batches_tracked += 1
        if clear_n:
            stats[:clear_n].zero_()
        return 0

    def tss_bn_fold(self, gamma, beta, running_mean, running_var, eps, scale, shift, C):
        scale.copy_(gamma.detach() / torch.sqrt(running_var + eps))
        shift.copy_(beta.detach() - running_mean * scale)
        return 0

    def tss_bn_apply(self, y, scale, shift, y2, scale2, shift2, res, z, M, C, ldy, ldy2, ldr, ldz, flags, dtype):
        v = y.float() * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if y2 is not None:
            v = v + y2.float() * scale2.view(1, -1, 1, 1) + shift2.view(1, -1, 1, 1)
        if res is not None:
            v = v + res.float()
        if flags & RELU:
            v = v.clamp_min(0)
        z.copy_(v)
        return 0

    @staticmethod
    def _g(dz, z, flags, y=None, mean=None, rstd=None, gamma=None, beta=None):
        g = dz.float()
        if flags & RELU:
            if z is None:      # mask recomputed from y with the forward's arithmetic
                v = lambda t: t.detach().view(1, -1, 1, 1)
                sc = v(gamma) * v(rstd)
                z = torch.addcmul(v(beta) - v(mean) * sc, y.float(), sc)
            g = g * (z.float() > 0)
        return g

    def tss_bn_bwd_reduce(self, dz, z, y, mean, rstd, gamma, beta, sums, M, C, lddz, ldz, ldy, flags, dtype):
        g = self._g(dz, z, flags, y, mean, rstd, gamma, beta)
        xh = (y.float() - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
        sums[:C] += g.sum(dim=(0, 2, 3))
        sums[C:] += (g * xh).sum(dim=(0, 2, 3))
        return 0

    def tss_bn_finalize_apply(self, stats, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum,
                              eps, mean, rstd, ticket, clear_n, y, res, z, M, C, ldy, ldr, ldz, flags, dtype):
        assert int(ticket) == 0
        scale, shift = torch.empty(C), torch.empty(C)
        self.tss_bn_finalize(stats, count, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
                             scale, shift, mean, rstd, C, clear_n)
        return self.tss_bn_apply(y, scale, shift, None, None, None, res, z, M, C, ldy, 0, ldr, ldz, flags, dtype)

    def tss_bn_bwd_apply(self, dz, z, y, mean, rstd, gamma, beta, sums, dy, dres, dgamma, dbeta, M, count, C, lddz, ldz,
                         ldy, lddy, lddres, flags, dtype):
        M = count if count > 0 else M
        g = self._g(dz, z, flags, y, mean, rstd, gamma, beta)
        v = lambda t: t.detach().view(1, -1, 1, 1)
        xh = (y.float() - v(mean)) * v(rstd)
        dy.copy_(v(gamma) * v(rstd) * (g - v(sums[:C]) / M - xh * v(sums[C:]) / M))
        if dres is not None:
            dres.copy_(g)
        if dbeta is not None:
            dbeta += sums[:C]
        if dgamma is not None:
            dgamma += sums[C:]
        return 0

    def tss_bn_bwd_onepass(self, dz, z, y, mean, rstd, gamma, beta, sums, dy, dres, dgamma, dbeta, M, C, lddz, ldz, ldy, lddy,
                           lddres, flags, sync, dtype):
        assert sync.dtype == torch.int32 and sync.numel() >= 4 and int(sync[2]) == 0
        self.tss_bn_bwd_reduce(dz, z, y, mean, rstd, gamma, beta, sums, M, C, lddz, ldz, ldy, flags, dtype)
        return self.tss_bn_bwd_apply(dz, z, y, mean, rstd, gamma, beta, sums, dy, dres, dgamma, dbeta, M, 0, C, lddz, ldz,
                                     ldy, lddy, lddres, flags, dtype)

    def tss_relu_bwd(self, dz, z, g, M, C, lddz, ldz, ldg, dtype):
        g.copy_(dz.float() * (z.float() > 0))
        return 0

    def tss_add(self, a, b, out, M, C, lda, ldb, ldo, dtype):
        out.copy_(a.float() + b.float())
        return 0

    def tss_copy_rows(self, src, dst, M, C, lds, ldd, dtype):
        dst.copy_(src)
        return 0

    def tss_cast_from_f32(self, src, dst, n, dtype):
        dst.copy_(src)
        return 0

    def tss_scale_inplace(self, x, s, n, dtype):
        x.copy_(x.float() * s)
        return 0

    # ---------------------------------------------------------------- pooling / resize
    def tss_adaptive_pool_fwd(self, x, out, N, H, W, C, bins, nbins, dtype):
        off = 0
        for b in bins.values:
            p = F.adaptive_avg_pool2d(x.float(), b).permute(0, 2, 3, 1).reshape(N * b * b, C)
            out[off:off + N * b * b].copy_(p)
            off += N * b * b
        return 0

    def tss_adaptive_pool_bwd(self, dout, dx, N, H, W, C, bins, nbins, accumulate, dtype):
        acc = dx.float().clone() if accumulate else torch.zeros(dx.shape)
        off = 0
        for b in bins.values:
            g = dout[off:off + N * b * b].float().view(N, b, b, C).permute(0, 3, 1, 2)
            xx = torch.zeros(N, C, H, W, requires_grad=True)
            with torch.enable_grad():
                F.adaptive_avg_pool2d(xx, b).backward(g)
            acc = acc + xx.grad
            off += N * b * b
        dx.copy_(acc)
        return 0

    def tss_bilinear_fwd(self, x, y, N, Hi, Wi, Ho, Wo, C, ldx, ldy, dtype):
        y.copy_(F.interpolate(x.float(), size=(Ho, Wo), mode='bilinear', align_corners=True))
        return 0

    def tss_bilinear_bwd(self, dy, dx, workspace, N, Hi, Wi, Ho, Wo, C, lddy, lddx, dtype):
        xx = torch.zeros(N, C, Hi, Wi, requires_grad=True)
        with torch.enable_grad():
            F.interpolate(xx, size=(Ho, Wo), mode='bilinear', align_corners=True).backward(dy.float())
        dx.copy_(xx.grad)
        return 0

    def tss_upsample_logits_fwd(self, x, y, N, Hi, Wi, Ho, Wo, C, ldx, dtype):
        y.copy_(F.interpolate(x.float(), size=(Ho, Wo), mode='bilinear', align_corners=True))
        return 0

    def tss_upsample_logits_bwd(self, dy, dx32, N, Hi, Wi, Ho, Wo, C, lddx, dtype):
        xx = torch.zeros(N, C, Hi, Wi, requires_grad=True)
        with torch.enable_grad():
            F.interpolate(xx, size=(Ho, Wo), mode='bilinear', align_corners=True).backward(dy.float())
        dx32[..., :C] += xx.grad.permute(0, 2, 3, 1)
        return 0

    def tss_bilinear_nchw_f32(self, x, y, NC, Hi, Wi, Ho, Wo):
        y.copy_(F.interpolate(x, size=(Ho, Wo), mode='bilinear', align_corners=True))
        return 0

    # ---------------------------------------------------------------- loss / metrics
    def tss_ce_count_valid(self, target, n, ignore_index, num_classes, nvalid):
        nvalid.fill_(int(((target != ignore_index) & (target >= 0) & (target < num_classes)).sum()))
        return 0

    @staticmethod
    def _ohem_weights(nll, ohem):
        cut, above, tie, wtie = [float(v) for v in ohem]
        return torch.where(nll > cut, torch.full_like(nll, above), torch.where(nll == tie, torch.full_like(nll, wtie), torch.zeros_like(nll)))

    def tss_ohem_workspace_bytes(self):
        return 16384

    def tss_ohem_select(self, pixel_loss, n, n_keep, thresh, workspace, loss, weights):
        v = pixel_loss.reshape(-1).float()
        srt, _ = torch.sort(v, descending=True)
        vk = srt[n_keep]
        if float(vk) > thresh:
            sel = v[v > thresh]
            loss.fill_(float(sel.mean()))
            weights.copy_(torch.tensor([thresh, 1.0 / sel.numel(), -1.0, 0.0]))
        else:
            loss.fill_(float(srt[:n_keep].mean()) if n_keep > 0 else float('nan'))
            gt, eq = int((v > vk).sum()), int((v == vk).sum())
            weights.copy_(torch.tensor([float(vk), 1.0 / max(n_keep, 1), float(vk), (n_keep - gt) / (eq * max(n_keep, 1)) if eq else 0.0]))
        return 0

    def tss_ce_fwd(self, logits, target, N, C, HW, ignore_index, nvalid, loss_sum, pixel_loss, dlogits, ohem, dtype):
        x = logits.float()
        logp = x.log_softmax(1)
        valid = target != ignore_index
        t = torch.where(valid, target, torch.zeros_like(target))
        nll = -logp.gather(1, t.unsqueeze(1)).squeeze(1) * valid
        if loss_sum is not None:
            loss_sum += nll.double().sum()
        if pixel_loss is not None:
            pixel_loss.copy_(nll)
        if dlogits is not None:
            p = logp.exp()
            p.scatter_add_(1, t.unsqueeze(1), -torch.ones_like(p[:, :1]))
            if ohem is not None:
                dlogits.copy_(p * (valid * self._ohem_weights(nll, ohem)).unsqueeze(1))
            else:
                nv = float(nvalid.item())
                dlogits.copy_(p * valid.unsqueeze(1) / nv if nv > 0 else torch.zeros_like(p))
        return 0

    def tss_upsample_ce_fwd(self, x, target, N, C, Hi, Wi, Ho, Wo, ldx, ignore_index, loss_sum, nvalid, pixel_loss,
                            dx32, lddx, ohem, dtype):
        xs = x.detach().float().clone().requires_grad_(True)
        with torch.enable_grad():
            logits = F.interpolate(xs, size=(Ho, Wo), mode='bilinear', align_corners=True)
            valid = (target != ignore_index) & (target >= 0) & (target < C)
            t = torch.where(valid, target, torch.zeros_like(target))
            nll = -logits.log_softmax(1).gather(1, t.unsqueeze(1)).squeeze(1) * valid
            total = nll.double().sum()
            plain = total.detach()
            if ohem is not None:
                total = (nll * self._ohem_weights(nll.detach(), ohem)).double().sum()
        if loss_sum is not None:
            loss_sum += plain
        nvalid += int(valid.sum())
        if pixel_loss is not None:
            pixel_loss.copy_(nll.detach())
        if dx32 is not None:
            (g,) = torch.autograd.grad(total, xs)
            dx32[..., :C] += g.permute(0, 2, 3, 1).float()
        return 0

    def tss_upsample_ce_finalize(self, loss_sum, nvalid, loss, dx32, dx, n, dtype):
        nv = float(nvalid.item())
        if loss is not None:
            loss.fill_(float(loss_sum.item()) / nv if nv > 0 else float('nan'))
        if dx is not None:
            dx.copy_(dx32 / nv if nv > 0 else dx32 * float('nan'))
        return 0

    def tss_ce_finalize(self, loss_sum, nvalid, loss):
        nv = float(nvalid.item())
        loss.fill_(float(loss_sum.item()) / nv if nv > 0 else float('nan'))
        return 0

    def tss_confusion_from_labels(self, pred, target, n, C, cm):
        m = (target >= 0) & (target < C) & (pred >= 0) & (pred < C)
        cm += torch.bincount(C * target[m] + pred[m], minlength=C * C).view(C, C)
        return 0

    def tss_confusion_from_logits(self, logits, target, N, C, HW, cm, pred_out, dtype):
        x = logits.float()
        pred = torch.where(torch.isnan(x), torch.full_like(x, math.inf), x).argmax(1)
        if pred_out is not None:
            pred_out.copy_(pred)
        return self.tss_confusion_from_labels(pred.reshape(-1), target.reshape(-1), pred.numel(), C, cm)

    def tss_adamw_step(self, p, g, m, v, n, hyper, grad_scale):
        lr, b1, b2, eps, wd = [float(hyper[i]) for i in range(5)]
        step = float(hyper[5]) + 1
        hyper[5] = step
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        hyper[6], hyper[7] = 1 / bc1, math.sqrt(bc2)
        gr = g * grad_scale
        p.mul_(1 - lr * wd)
        m.lerp_(gr, 1 - b1)
        v.mul_(b2).addcmul_(gr, gr, value=1 - b2)
        p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
        return 0
