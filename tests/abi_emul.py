"""CPU emulation of the libb200gan C ABI — TEST INFRASTRUCTURE ONLY.

Each method restates, with plain torch CPU ops, the contract a C entry point documents in include/b200gan.h.  It is
(1) the executable spec the `-m gpu` op tests compare every kernel against and (2) a stand-in that lets the
`-m "not gpu"` suite exercise the host-side wiring (autograd formulas, descriptors, packing plans, module surface)
without a GPU.  Nothing in the product imports this file.
"""
import torch
import torch.nn.functional as F

MODE_PLAIN, MODE_AFFINE, MODE_CBN, MODE_SPADE = 0, 1, 2, 3


def _flat(t):
    return t.reshape(-1) if t.is_contiguous() else None


def _storage_view(t: torch.Tensor, offset: int, size, stride):
    """View into the underlying storage of `t` starting `offset` elements after t's own first element."""
    return torch.as_strided(t, size, stride, t.storage_offset() + offset)


def _tf32(x: torch.Tensor) -> torch.Tensor:
    """round an fp32 tensor to tf32 (10-bit mantissa, nearest) — what the TFLOAT32 tensor maps do on the way to shared memory"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class EmulKernels:
    def __init__(self):
        self.launches = 0

    def launch_count(self):
        return self.launches

    def version(self):
        return 100

    def conv_tc_ntile(self, cout):
        return 128 if cout >= 128 else (64 if cout >= 64 else 16)

    def bn_chunks(self, rows, C):
        return 4

    # ---- crops ----------------------------------------------------------------------------------------
    @staticmethod
    def _coords(boxes, w, lo, hi, size):
        S = w.numel() // 2
        sw, ew = w[:S], w[S:]
        a = 2 * boxes[:, lo] - 1
        b = 2 * boxes[:, hi] - 1
        g = sw[None] * a[:, None] + ew[None] * b[:, None]
        return ((g + 1) * size - 1) / 2

    def crop_taps(self, boxes, wx, wy, H, W, HH, WW):
        ix = self._coords(boxes, wx, 0, 2, W)
        iy = self._coords(boxes, wy, 1, 3, H)
        ix0, iy0 = torch.floor(ix), torch.floor(iy)
        return ix0.to(torch.int32), iy0.to(torch.int32), ix - ix0, iy - iy0

    def _crop_matrices(self, boxes, wx, wy, H, W):
        ix0, iy0, fx, fy = self.crop_taps(boxes, wx, wy, H, W, wy.numel() // 2, wx.numel() // 2)
        B = boxes.shape[0]

        def mat(i0, f, size):
            S = i0.shape[1]
            M = torch.zeros(B, S, size + 2)
            idx = (i0.long() + 1).clamp(0, size + 1)           # shift by one so -1 lands in a discarded column
            ok0 = (i0 >= 0) & (i0 < size)
            ok1 = (i0 + 1 >= 0) & (i0 + 1 < size)
            M.scatter_add_(2, idx.unsqueeze(-1), ((1 - f) * ok0).unsqueeze(-1))
            idx1 = (i0.long() + 2).clamp(0, size + 1)
            M.scatter_add_(2, idx1.unsqueeze(-1), (f * ok1).unsqueeze(-1))
            return M[:, :, 1:size + 1]

        return mat(iy0, fy, H), mat(ix0, fx, W)                 # (B,HH,H), (B,WW,W)

    def crop_fwd(self, feats, boxes, box_to_img, wx, wy, HH, WW):
        self.launches += 1
        N, C, H, W = feats.shape
        My, Mx = self._crop_matrices(boxes, wx, wy, H, W)
        src = feats[box_to_img.long()]
        return torch.einsum("bih,bchw,bjw->bcij", My, src, Mx)

    def crop_bwd(self, dcrops, boxes, img_box_start, box_order, wx, wy, N, H, W):
        self.launches += 2
        B, C = dcrops.shape[:2]
        My, Mx = self._crop_matrices(boxes, wx, wy, H, W)
        d = torch.einsum("bih,bcij,bjw->bchw", My, dcrops, Mx)
        out = torch.zeros(N, C, H, W)
        start = img_box_start.tolist()
        order = box_order.long()
        for n in range(N):
            ids = order[start[n]:start[n + 1]]
            if ids.numel():
                out[n] = d[ids].sum(0)
        return out

    # ---- gather GEMMs -------------------------------------------------------------------------------------
    @staticmethod
    def _gather(d, inp):
        """A[m, tap, c] per the descriptor; returns (M, T, Cin) and the output row offsets."""
        B, Qh, Qw = d.B, d.Qh, d.Qw
        Hp = ((d.Hi - 1) >> d.up_shift) + 1
        Wp = ((d.Wi - 1) >> d.up_shift) + 1
        x = _storage_view(inp, 0, (B, Hp, Wp, d.Cin), (d.in_sn, d.in_sh, d.in_sw, d.in_sc))
        qy = torch.arange(Qh)
        qx = torch.arange(Qw)
        cols = []
        for ty in range(d.Th):
            for tx in range(d.Tw):
                iy = qy * d.in_sy + ty * d.tap_sy + d.tap_oy
                ix = qx * d.in_sx + tx * d.tap_sx + d.tap_ox
                vy = (iy >= 0) & (iy < d.Hi)
                vx = (ix >= 0) & (ix < d.Wi)
                g = x[:, (iy.clamp(0, d.Hi - 1) >> d.up_shift)][:, :, (ix.clamp(0, d.Wi - 1) >> d.up_shift)]
                g = g * (vy[:, None] & vx[None, :]).to(g.dtype)[None, :, :, None]
                cols.append(g.reshape(B * Qh * Qw, d.Cin))
        return torch.stack(cols, dim=1)

    @staticmethod
    def _out_index(d):
        qy = torch.arange(d.Qh)
        qx = torch.arange(d.Qw)
        oy = qy * d.out_sy + d.out_oy
        ox = qx * d.out_sx + d.out_ox
        ok = ((oy >= 0) & (oy < d.Ho))[:, None] & ((ox >= 0) & (ox < d.Wo))[None, :]
        off = oy[:, None] * d.out_sh + ox[None, :] * d.out_sw
        n = torch.arange(d.B) * d.out_sn
        off = (n[:, None, None] + off[None]).reshape(-1)
        ok = ok[None].expand(d.B, -1, -1).reshape(-1)
        return off, ok

    def cast_bf16(self, x):
        self.launches += 1
        return x.to(torch.bfloat16)

    def split_bf16(self, x):
        self.launches += 1
        hi = x.to(torch.bfloat16)
        return hi, (x - hi.float()).to(torch.bfloat16)

    def im2col_pack(self, x, x_strides, N, Hx, Wx, Cx, kh, kw, stride, pad, Hy, Wy, Kp, flip=False):
        self.launches += 1
        sn, sh, sw, sc = x_strides
        xv = torch.as_strided(x, (N, Cx, Hx, Wx), (sn, sc, sh, sw)).float()
        out = torch.zeros((N, Hy, Wy, Kp), dtype=torch.float32)
        qy = torch.arange(Hy)[:, None]
        qx = torch.arange(Wy)[None, :]
        for ky in range(kh):
            for kx in range(kw):
                iy = qy * stride + (pad - ky if flip else ky - pad)
                ix = qx * stride + (pad - kx if flip else kx - pad)
                ok = (iy >= 0) & (iy < Hx) & (ix >= 0) & (ix < Wx)                            # (Hy, Wy)
                g = xv[:, :, iy.clamp(0, Hx - 1).expand(Hy, Wy), ix.clamp(0, Wx - 1).expand(Hy, Wy)]   # (N, Cx, Hy, Wy)
                k0 = (ky * kw + kx) * Cx
                out[..., k0:k0 + Cx] = torch.where(ok, g, torch.zeros(())).permute(0, 2, 3, 1)
        return out.reshape(N * Hy * Wy, Kp).to(torch.bfloat16)

    def rowsum(self, x2d):
        self.launches += 1
        return x2d.double().sum(1).float()

    def conv_stats_ok(self, d, tc):
        """b200_conv_tc_stats_ok, approximated: a tcgen05 launch with a 64+ wide N tile and no fused ReLU mask"""
        return bool(int(tc)) and d.Cout >= 64 and not d.relu_mask

    def conv_stats_buffer(self, d, device):
        M = d.B * d.Qh * d.Qw
        nt = self.conv_tc_ntile(d.Cout)
        return torch.zeros(((M + 127) // 128 * 4, (d.Cout + nt - 1) // nt * nt, 2))

    def bn_stats_slabs(self, col_stats, rows, C, running_mean, running_var, momentum, groups=1):
        self.launches += 1
        rpg = rows // groups
        assert rpg % 32 == 0
        cs = col_stats[:rows // 32, :C].double().view(groups, rpg // 32, C, 2).sum(1)
        mean = cs[..., 0] / rpg
        var = (cs[..., 1] / rpg - mean * mean).clamp(min=0)
        if running_mean is not None:
            for g in range(groups):
                unb = var[g] * rpg / (rpg - 1) if rpg > 1 else var[g]
                running_mean.mul_(1 - momentum).add_(momentum * mean[g].float())
                running_var.mul_(1 - momentum).add_(momentum * unb.float())
        return mean.float(), var.float()

    def conv_gemm(self, d, inp, wmat, bias, scale, out, tc, mask=None, stats=None):
        self.launches += 1
        tc = int(tc)
        assert mask is None or (tc and mask.shape == out.shape and mask.dtype == out.dtype)
        assert inp.dtype == torch.bfloat16 or tc != 1, "the tcgen05 bf16 path takes bf16 operands"
        assert tc != 2 or (inp.dtype == torch.float32 and wmat.dtype == torch.float32 and out.dtype == torch.float32)
        A = self._gather(d, inp).float()
        K = d.Th * d.Tw * d.Cin
        Wm = wmat.float()[:d.Cout, :K]
        A2 = A.reshape(A.shape[0], K)
        if tc == 1:
            A2 = A2.to(torch.bfloat16).float()
        elif tc == 2:
            A2, Wm = _tf32(A2), _tf32(Wm)
        y = A2 @ Wm.t()
        if scale is not None:
            sr = getattr(d, "scale_rows", 0)
            if sr > 0:
                y = y * scale.reshape(-1)[(torch.arange(y.shape[0]) // sr)][:, None]
            else:
                y = y * scale.reshape(-1)[0]
        if bias is not None:
            y = y + bias[None]
        if d.relu:
            y = F.relu(y)
        off, ok = self._out_index(d)
        flat = torch.as_strided(out, (out.untyped_storage().nbytes() // out.element_size() - out.storage_offset(),), (1,),
                                out.storage_offset())
        co = torch.arange(d.Cout) * d.out_sc
        idx = (off[ok][:, None] + co[None]).reshape(-1)
        vals = y[ok].reshape(-1)
        if mask is not None:
            mflat = torch.as_strided(mask, flat.shape, (1,), mask.storage_offset())
            vals = torch.where(mflat[idx].float() > 0, vals, torch.zeros(()))
        flat[idx] = vals.to(out.dtype)
        if stats is not None:      # b200_conv_desc.col_stats: (sum, sum^2) of the STORED values per 32-row slab and channel
            assert int(tc) and mask is None
            M = y.shape[0]
            q = torch.where(ok[:, None], y.to(out.dtype).float(), torch.zeros(()))
            pad = stats.shape[0] * 32 - M
            q = torch.cat([q, torch.zeros(pad, q.shape[1])]) if pad else q
            q = q.view(stats.shape[0], 32, -1).double()
            stats.zero_()
            stats[:, :d.Cout, 0] = q.sum(1).float()
            stats[:, :d.Cout, 1] = (q * q).sum(1).float()

    def wgrad_gemm(self, d, P, G, ws, splits, tc):
        self.launches += 1
        tc = int(tc)
        assert (P.dtype == torch.bfloat16 and G.dtype == torch.bfloat16) or tc != 1
        assert tc != 2 or (P.dtype == torch.float32 and G.dtype == torch.float32)
        A = self._gather(d, G).float()                           # (Q, T, Cin)
        off, ok = self._out_index(d)
        flatP = torch.as_strided(P, (P.untyped_storage().nbytes() // P.element_size() - P.storage_offset(),), (1,),
                                 P.storage_offset())
        co = torch.arange(d.Cout) * d.out_sc
        Pm = flatP[(off.clamp(min=0)[:, None] + co[None])].float() * ok[:, None].float()
        if tc == 1:
            A = A.to(torch.bfloat16).float()
            Pm = Pm.to(torch.bfloat16).float()
        elif tc == 2:
            A, Pm = _tf32(A), _tf32(Pm)
        K = d.Th * d.Tw * d.Cin
        A2 = A.reshape(A.shape[0], K)
        wsv = ws.view(splits, d.Cout, K)
        # the kernels split the pixel range into `splits` chunks of rows_per_split = ceil(Q / splits) rounded up to 64
        Q = A2.shape[0]
        rps = max(64, (-(-Q // splits) + 63) // 64 * 64)
        for z in range(splits):
            a, b = z * rps, min(Q, (z + 1) * rps)
            wsv[z] = Pm[a:b].t() @ A2[a:b] if b > a else 0.0

    def wgrad_reduce(self, ws, splits, M, Th, Tw, C, dst, dst_offset, s_m, s_ty, s_tx, s_c, scale=None, accumulate=False,
                     ws_row_offset=0, ws_rows=None):
        self.launches += 1
        rows = M if ws_rows is None else ws_rows
        K = Th * Tw * C
        R = ws.view(splits, rows, K)[:, ws_row_offset:ws_row_offset + M].sum(0).view(M, Th, Tw, C)
        if scale is not None:
            R = R * scale.reshape(-1)[0]
        view = _storage_view(dst, dst_offset, (M, Th, Tw, C), (s_m, s_ty, s_tx, s_c))
        if accumulate:
            view += R
        else:
            view.copy_(R)

    def pack_weight(self, src, src_offset, dst, dst_row_offset, bf16, M, Mpad, Th, Tw, C, ldw, s_m, s_ky, s_kx, s_c,
                    ky0=0, kx0=0, kstep=1, C_dst=0, c_off=0):
        self.launches += 1
        self._pack(src, src_offset, dst, dst_row_offset, M, Mpad, Th, Tw, C, ldw, s_m, s_ky, s_kx, s_c, ky0, kx0, kstep,
                   C_dst, c_off, valid_only=False)

    @staticmethod
    def _pack(src, src_offset, dst, dst_row_offset, M, Mpad, Th, Tw, C, ldw, s_m, s_ky, s_kx, s_c, ky0, kx0, kstep, C_dst,
              c_off, valid_only):
        v = _storage_view(src.detach(), src_offset + ky0 * s_ky + kx0 * s_kx, (M, Th, Tw, C),
                          (s_m, s_ky * kstep, s_kx * kstep, s_c))
        full = C_dst <= 0 or (C_dst == C and c_off == 0)
        if C_dst <= 0:
            C_dst, c_off = C, 0
        with torch.no_grad():
            if full and not valid_only:          # whole Mpad x ldw block, zero padded
                out = torch.zeros(Mpad, ldw)
                out[:M, :Th * Tw * C] = v.reshape(M, -1)
                dst[dst_row_offset:dst_row_offset + Mpad] = out.to(dst.dtype)
            else:                                # only the valid elements (of a channel slice)
                blk = dst[dst_row_offset:dst_row_offset + M, :Th * Tw * C_dst].view(M, Th * Tw, C_dst)
                blk[:, :, c_off:c_off + C] = v.reshape(M, Th * Tw, C).to(dst.dtype)

    # ---- fused step arithmetic (loss.cu): values through torch's own functional ops, gradients through autograd ------
    @staticmethod
    def _loss_finish(loss_terms, xs, partials, counts, slot):
        """loss_terms: list of scalar tensors (one per slot, starting at `slot`); returns the gradients of their sum"""
        grads = torch.autograd.grad(sum(loss_terms), xs, allow_unused=True)
        for i, t in enumerate(loss_terms):
            partials[slot + i].zero_()
            partials[slot + i, 0] = float(t)
            counts[slot + i] = 1
        return [g if g is not None else torch.zeros_like(x) for g, x in zip(grads, xs)]

    def loss_bce_groups(self, x, n, groups, split_group, target, weight, scale, partials, counts, slot):
        self.launches += 1
        xv = x.detach().clone().requires_grad_(True)
        per = F.binary_cross_entropy_with_logits(xv.view(groups, n), target.view(groups, 1).expand(groups, n), reduction="none").mean(1)
        terms = [scale * (weight[:split_group] * per[:split_group]).sum()]
        if split_group < groups:
            terms.append(scale * (weight[split_group:] * per[split_group:]).sum())
        return self._loss_finish(terms, [xv], partials, counts, slot)[0].view(x.shape)

    def loss_ce_groups(self, x, label, n, groups, weight, scale, partials, counts, slot):
        self.launches += 1
        xv = x.detach().clone().requires_grad_(True)
        per = F.cross_entropy(xv, label.repeat(groups), reduction="none").view(groups, n).mean(1)
        return self._loss_finish([scale * (weight * per).sum()], [xv], partials, counts, slot)[0]

    def loss_bce_pw_rows(self, x, t, sel, n, groups, n_sel, pos_weight, weight, scale, partials, counts, slot):
        self.launches += 1
        xv = x.detach().clone().requires_grad_(True)
        idx = sel.nonzero().view(-1)
        if idx.numel() == 0:
            partials[slot].zero_(); counts[slot] = 1
            return torch.zeros_like(x)
        term = 0.0
        for g in range(groups):
            term = term + weight[g] * F.binary_cross_entropy_with_logits(xv[g * n:(g + 1) * n].index_select(0, idx),
                                                                         t.index_select(0, idx), pos_weight=pos_weight)
        return self._loss_finish([scale * term], [xv], partials, counts, slot)[0]

    def loss_l1_rows(self, a, b, N, L, b_stride_n, mask, denom, scale, partials, counts, slot):
        self.launches += 1
        av = a.detach().clone().requires_grad_(True)
        bb = b.reshape(1, L).expand(N, L) if b_stride_n == 0 else b.reshape(N, L)
        per = (av.reshape(N, L) - bb).abs().mean(1)
        m = mask if mask is not None else torch.ones(N)
        return self._loss_finish([scale * (m * per).sum() / denom], [av], partials, counts, slot)[0].view(a.shape)

    def loss_kl(self, mu, logvar, scale, partials, counts, slot):
        self.launches += 1
        m, lv = mu.detach().clone().requires_grad_(True), logvar.detach().clone().requires_grad_(True)
        term = scale * (-0.5) * torch.sum(1 + lv - m.pow(2) - lv.exp())
        g = self._loss_finish([term], [m, lv], partials, counts, slot)
        return g[0], g[1]

    def loss_total(self, partials, counts, n_terms):
        self.launches += 1
        terms = torch.zeros(n_terms + 1)
        for s_ in range(n_terms):
            terms[s_] = float(partials[s_, :int(counts[s_])].sum())
        terms[n_terms] = float(terms[:n_terms].double().sum())
        return terms

    def pack_table(self, recipes, device):
        return None, list(recipes), len(recipes), 0

    def pack_weight_multi(self, table_dev, n_entries, total_chunks):
        """b200_pack_weight_multi: every entry's valid elements, padding untouched"""
        self.launches += 1
        for r in table_dev:
            self._pack(r.src, r.src_offset, r.dst, r.dst_row_offset, r.M, r.Mpad, r.Th, r.Tw, r.C, r.ldw, r.s_m, r.s_ky,
                       r.s_kx, r.s_c, r.ky0, r.kx0, r.kstep, r.C_dst, r.c_off, valid_only=True)

    # ---- normalisation --------------------------------------------------------------------------------------
    def bn_stats(self, x2d, running_mean, running_var, momentum, groups=1):
        self.launches += 2
        x2d = x2d.float()
        rows, C = x2d.shape
        rpg = rows // groups
        xd = x2d.double().view(groups, rpg, C)
        mean = xd.mean(1)
        var = ((xd * xd).mean(1) - mean * mean).clamp(min=0)
        if running_mean is not None:
            for g in range(groups):
                unb = var[g] * rpg / (rpg - 1) if rpg > 1 else var[g]
                running_mean.mul_(1 - momentum).add_(momentum * mean[g].float())
                running_var.mul_(1 - momentum).add_(momentum * unb.float())
        return mean.float(), var.float()

    @staticmethod
    def _g_b(mode, gamma, beta, idx, rows_per_seg, rows, C):
        if mode == MODE_PLAIN:
            return None, None
        if mode == MODE_AFFINE:
            return gamma[None], beta[None]
        if mode == MODE_CBN:
            t = gamma[idx.long()].repeat_interleave(rows_per_seg, dim=0)
            return t[:, :C], t[:, C:]
        return 1 + gamma[:, :C], gamma[:, C:]

    def norm_fwd(self, x2d, mean, var, eps, mode, gamma, beta, idx, rows_per_seg, residual, relu, groups=1):
        self.launches += 1
        dt = x2d.dtype
        assert residual is None or residual.dtype == dt
        assert mode != MODE_SPADE or gamma.dtype == dt
        x2d, gamma = x2d.float(), (gamma.float() if gamma is not None else None)
        residual = residual.float() if residual is not None else None
        rows, C = x2d.shape
        rpg = rows // groups
        mean_r = mean.view(groups, C).repeat_interleave(rpg, dim=0)
        rstd_r = (1.0 / torch.sqrt(var.view(groups, C) + eps)).repeat_interleave(rpg, dim=0)
        xh = (x2d - mean_r) * rstd_r
        g, b = self._g_b(mode, gamma, beta, idx, rows_per_seg, rows, C)
        y = xh if g is None else xh * g + b
        if residual is not None:
            y = y + residual
        return (F.relu(y) if relu else y).to(dt)

    def norm_bwd(self, dy, x2d, y, mean, var, eps, mode, gamma, idx, rows_per_seg, relu, num_classes, groups=1, sync=None):
        self.launches += 4
        dt = x2d.dtype
        assert dy.dtype == dt and (y is None or y.dtype == dt) and (mode != MODE_SPADE or gamma.dtype == dt)
        dy, x2d, y = dy.float(), x2d.float(), (y.float() if y is not None else None)
        gamma = gamma.float() if gamma is not None else None
        rows, C = x2d.shape
        rpg = rows // groups
        mean_r = mean.view(groups, C).repeat_interleave(rpg, dim=0)
        rstd_r = (1.0 / torch.sqrt(var.view(groups, C) + eps)).repeat_interleave(rpg, dim=0)
        xh = (x2d - mean_r) * rstd_r
        if relu == 2:        # recomputed mask (conditional batch norm): gamma * xhat + beta > 0
            assert mode == MODE_CBN and y is None
            gm2, bm2 = self._g_b(mode, gamma, None, idx, rows_per_seg, rows, C)
            g = dy * ((gm2 * xh + bm2) > 0).to(dy.dtype)
        else:
            g = dy * (y > 0).to(dy.dtype) if relu else dy
        gm, _ = self._g_b(mode, gamma, gamma if mode == MODE_AFFINE else None, idx, rows_per_seg, rows, C)
        dxh = g if gm is None else g * gm
        s1 = dxh.double().view(groups, rpg, C).sum(1).float()
        s2 = (dxh.double() * xh.double()).view(groups, rpg, C).sum(1).float()
        if sync is not None:          # b200_norm_bwd_finalize's output layout: [group][channel][2]
            sg = sync(torch.stack([s1, s2], dim=2).reshape(-1).contiguous()).view(groups, C, 2)
            s1, s2 = sg[:, :, 0], sg[:, :, 1]
        s1, s2 = s1.repeat_interleave(rpg, dim=0), s2.repeat_interleave(rpg, dim=0)
        dx = rstd_r * (dxh - s1 / rpg - xh * s2 / rpg)
        dgamma = dbeta = dtable = dgb = None
        if mode == MODE_AFFINE:
            dgamma = (g.double() * xh.double()).sum(0).float()
            dbeta = g.double().sum(0).float()
        elif mode == MODE_CBN:
            nseg = rows // rows_per_seg
            a = (g * xh).view(nseg, rows_per_seg, C).sum(1)
            b = g.view(nseg, rows_per_seg, C).sum(1)
            dtable = torch.zeros(num_classes, 2 * C)
            dtable.index_add_(0, idx.long(), torch.cat([a, b], dim=1))
        elif mode == MODE_SPADE:
            dgb = torch.cat([g * xh, g], dim=1).to(dt)
        return dx.to(dt), dgamma, dbeta, dtable, dgb

    # ---- elementwise ----------------------------------------------------------------------------------------
    def relu_fwd(self, x):
        self.launches += 1
        return F.relu(x)

    def relu_bwd(self, dy, y):
        self.launches += 1
        assert dy.dtype == y.dtype
        return dy * (y > 0).to(dy.dtype)

    def add(self, a, b, out=None):
        self.launches += 1
        assert a.dtype == b.dtype
        r = (a.float() + b.float()).to(a.dtype)
        if out is None:
            return r
        out.copy_(r)
        return out

    def pool_fwd(self, x, N, H, W, C, f, scale):
        self.launches += 1
        return (x.float().reshape(N, H // f, f, W // f, f, C).sum(dim=(2, 4)) * scale).to(x.dtype)

    def pool_add_fwd(self, a, b, N, H, W, C, f, scale, relu=False):
        self.launches += 1
        y = (a.float() + b.float()).reshape(N, H // f, f, W // f, f, C).sum(dim=(2, 4)) * scale
        return (F.relu(y) if relu else y).to(a.dtype)

    def add_relu(self, a, b):
        self.launches += 1
        return F.relu(a.float() + b.float()).to(a.dtype)

    def unpool_masked_fwd(self, x, mask, N, H, W, C, f, scale):
        self.launches += 1
        v = (x.float() * (mask.float() > 0)).reshape(N, H, 1, W, 1, C).expand(N, H, f, W, f, C)
        return (v * scale).reshape(N, H * f, W * f, C).to(x.dtype)

    def unpool_fwd(self, x, N, H, W, C, f, scale):
        self.launches += 1
        v = x.float().reshape(N, H, 1, W, 1, C).expand(N, H, f, W, f, C)
        return (v * scale).reshape(N, H * f, W * f, C).contiguous().to(x.dtype)

    def concat_fwd(self, a, Ca, a_div, b, Cb, b_div, rows):
        self.launches += 1
        assert a.dtype == b.dtype
        aa = a.reshape(-1, Ca).repeat_interleave(a_div, dim=0)
        bb = b.reshape(-1, Cb).repeat_interleave(b_div, dim=0)
        return torch.cat([aa, bb], dim=1)

    def concat_bwd(self, dout, Ca, a_div, Cb, b_div, rows, need_a=True, need_b=True):
        self.launches += 1
        d = dout.float().reshape(rows, Ca + Cb)
        da = d[:, :Ca].reshape(rows // a_div, a_div, Ca).sum(1).to(dout.dtype) if need_a else None
        db = d[:, Ca:].reshape(rows // b_div, b_div, Cb).sum(1).to(dout.dtype) if need_b else None
        return da, db

    def gather_rows(self, table, idx):
        self.launches += 1
        return table[idx.long()]

    def scatter_rows(self, dout, idx, num_classes):
        self.launches += 1
        out = torch.zeros(num_classes, dout.shape[1])
        out.index_add_(0, idx.long(), dout)
        return out

    def permute_rows(self, x, src_row, rowlen):
        self.launches += 1
        src = src_row.long()
        xv = x.reshape(-1, rowlen)
        out = xv[src.clamp(min=0)] * (src >= 0).to(x.dtype)[:, None]
        return out

    def mask_outer_fwd(self, v, mask, O, H, W, C, out_dtype=torch.float32):
        self.launches += 1
        out = torch.zeros(O, H + 2, W + 2, C)
        out[:, 1:H + 1, 1:W + 1] = mask.reshape(O, H, W, 1) * v.reshape(O, 1, 1, C)
        return out.to(out_dtype)

    def mask_outer_bwd(self, dout, mask, O, H, W, C):
        self.launches += 1
        return (dout.float()[:, 1:H + 1, 1:W + 1] * mask.reshape(O, H, W, 1)).sum(dim=(1, 2))

    def lstm_gates_fwd(self, pre_x, pre_h, c_prev, rows, hid, gates=None, c_out=None, h_out=None):
        self.launches += 1
        dt = pre_x.dtype
        assert (pre_h is None or pre_h.dtype == dt) and (h_out is None or h_out.dtype == dt)
        pre = pre_x.float().reshape(rows, 4 * hid)
        if pre_h is not None:
            pre = pre + pre_h.float().reshape(rows, 4 * hid)
        i, f, o, g = torch.split(pre, hid, dim=1)
        i, f, o, g = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(g)
        cp = c_prev.reshape(rows, hid) if c_prev is not None else 0
        cn = f * cp + i * g
        hn = o * torch.tanh(cn)
        gt = torch.cat([i, f, o, g], dim=1)
        if gates is None:
            return gt, cn, hn.to(dt)
        gates.copy_(gt.view(gates.shape))
        c_out.copy_(cn.view(c_out.shape))
        h_out.copy_(hn.view(h_out.shape))
        return gates, c_out, h_out

    def lstm_gates_bwd(self, dh, dc_next, gates, c_prev, c_out, rows, hid, dpre=None, dc_prev=None):
        self.launches += 1
        i, f, o, g = torch.split(gates.reshape(rows, 4 * hid), hid, dim=1)
        tc = torch.tanh(c_out.reshape(rows, hid))
        assert dpre is None or dpre.dtype == dh.dtype
        dhv = dh.float().reshape(rows, hid)
        dc = dhv * o * (1 - tc * tc)
        if dc_next is not None:
            dc = dc + dc_next.reshape(rows, hid)
        cp = c_prev.reshape(rows, hid) if c_prev is not None else torch.zeros(rows, hid)
        dp = torch.cat([dc * g * i * (1 - i), dc * cp * f * (1 - f), dhv * tc * o * (1 - o), dc * i * (1 - g * g)], dim=1)
        dcp = dc * f
        if dpre is None:
            return dp.to(dh.dtype), dcp
        dpre.copy_(dp.view(dpre.shape))
        dc_prev.copy_(dcp.view(dc_prev.shape))
        return dpre, dc_prev

    def reparam_fwd(self, mu, logvar, eps):
        self.launches += 1
        return eps * torch.exp(logvar * 0.5) + mu

    def reparam_bwd(self, dz, logvar, eps):
        self.launches += 1
        return dz.clone(), dz * eps * 0.5 * torch.exp(logvar * 0.5)

    def transpose(self, x, B, R, C):
        self.launches += 1
        return x.reshape(B, R, C).transpose(1, 2).contiguous()

    def colsum(self, x2d):
        self.launches += 2
        return x2d.double().sum(0).float()

    # ---- spectral norm ------------------------------------------------------------------------------------------
    def sn_power_iter(self, W, h, w, u, v, do_iter, eps, inv_out=None, sigma_out=None):
        self.launches += 4
        Wm = W.detach().reshape(h, w)
        if do_iter:
            v.copy_(F.normalize(torch.mv(Wm.t(), u), dim=0, eps=eps))
            wv = torch.mv(Wm, v)
            u.copy_(F.normalize(wv, dim=0, eps=eps))
        else:
            wv = torch.mv(Wm, v)
        sigma = torch.dot(u, wv)
        if inv_out is None:
            inv_out = torch.empty(1)
        inv_out.copy_((1.0 / sigma).reshape(1))
        if sigma_out is not None:
            sigma_out.copy_(sigma.reshape(1))
        return inv_out

    def sn_grad(self, g, W, u, v, inv_sigma, h, w, dW=None, accumulate=False):
        self.launches += 3
        inv = inv_sigma.reshape(-1)[0]
        dot = (g.double() * W.detach().double()).sum().float()
        val = (g.reshape(h, w) * inv - dot * inv * inv * torch.outer(u, v)).reshape(W.shape)
        if dW is None:
            return val
        if accumulate:
            dW += val
        else:
            dW.copy_(val)
        return dW

    def fold_pool_weight(self, w, w4):
        """b200_fold_pool_weight: W4[f][a][b] = 0.25 * sum_{i,j in {0,1}} W[f][a-i][b-j]"""
        self.launches += 1
        kh, kw = w.shape[-2], w.shape[-1]
        acc = torch.zeros_like(w4)
        for i in (0, 1):
            for j in (0, 1):
                acc[..., i:i + kh, j:j + kw] += w.detach()
        w4.copy_(0.25 * acc)
        return w4

    def sn_wgrad_finish(self, ws, groups, spg, Cy, T, Cx, W, u_hist, v_hist, inv, dW, pooled_taps=None):
        self.launches += 2
        if pooled_taps is not None:
            # b200_sn_wgrad_finish_pooled: partials of the folded convolution, transposed fold while summing
            th, tw = pooled_taps
            T = th * tw
            p4 = ws.view(groups, spg, Cy, th + 1, tw + 1, Cx).sum(1)
            part = 0.25 * ((p4[:, :, :-1, :-1] + p4[:, :, :-1, 1:]) + (p4[:, :, 1:, :-1] + p4[:, :, 1:, 1:]))
            part = part.reshape(groups, Cy, T, Cx)
        else:
            part = ws.view(groups, spg, Cy, T, Cx).sum(1)                 # (groups, Cy, tap, Cx)
        G = part.permute(0, 1, 3, 2).reshape(groups, Cy, Cx * T)           # parameter layout (Cy, Cx, T)
        Wm = W.detach().reshape(Cy, Cx * T)
        out = torch.zeros(Cy, Cx * T)
        for g in range(groups):
            dot = (G[g].double() * Wm.double()).sum().float()
            out += G[g] * inv[g] - dot * inv[g] * inv[g] * torch.outer(u_hist[g], v_hist[g])
        dW.copy_(out.reshape(dW.shape))
        return dW

    def sn_table(self, layers, stage, ws, iters):
        return None

    def sn_power_iter_multi(self, table, layers, stage, ws, iters, do_iter, eps):
        self.launches += 4 * iters
        for (W, u, v, h, w, so, wo) in layers:
            for it in range(iters):
                inv = stage[so + it:so + it + 1]
                self.launches -= 4
                self.sn_power_iter(W, h, w, u, v, do_iter, eps, inv_out=inv)
                stage[so + iters + it * h:so + iters + (it + 1) * h] = u
                stage[so + iters + iters * h + it * w:so + iters + iters * h + (it + 1) * w] = v

    def adam_multi(self, table, n_entries, step, lr, beta1, beta2, eps):
        raise NotImplementedError("adam_multi takes raw device pointers; the CPU wiring tests use torch.optim.Adam")

    def copy_into(self, dst, dst_row, src):
        self.launches += 1
        dst[dst_row:dst_row + src.shape[0]].copy_(src)
