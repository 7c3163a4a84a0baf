"""Pure-torch stand-in for distegnn_b200.backend.CudaBackend — TEST INFRASTRUCTURE.

It restates, stage by stage, what each C-ABI kernel computes (same decomposition: per-node P/Q/Hn
split of the first MLP layers, SUMS instead of means, packed vsum buffer) so that
  * the host-side orchestration in FastEGNN.forward (weight packing, flags, buffer reuse, the
    one-all-reduce-per-layer protocol) can be checked against the oracle on a machine without a GPU,
    including world_size=2 under gloo;
  * on the GPU each kernel can be compared with its stage here in isolation.
It is never imported by the package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from distegnn_b200 import _lib

H = 64


def _fields(lp, A, C, Na):
    offs, total = _lib.param_layout(A, C, Na)
    sizes = dict(E_W1A=(H, H), E_W1B=(H, H), E_W1R=(H,), E_W1E=(A, H), E_B1=(H,), E_W2=(H, H), E_B2=(H,),
                 E_WC=(H, H), E_BC=(H,), E_W3=(H,), V_W1H=(H, H), V_W1V=(H, H), V_W1R=(H,), V_W1M=(C, H),
                 V_B1=(H,), V_W2=(H, H), V_B2=(H,), V_WXV=(H, H), V_BXV=(H,), V_W3XV=(H,), V_WX=(H, H),
                 V_BX=(H,), V_W3X=(H,), L_W=(H, H), L_B=(H,), L_W3=(H,), L_B3=(1,), N_W1=(3 * H + Na, H),
                 N_B1=(H,), N_W2=(H, H), N_B2=(H,), M_W1=(2 * H, H), M_B1=(H,), M_W2=(H, H), M_B2=(H,))
    out = {}
    for k, shp in sizes.items():
        n = 1
        for s in shp:
            n *= s
        out[k] = lp[offs[k]:offs[k] + n].reshape(shp)
    return out


class ShadowBackend:
    name = "torch-shadow (tests only)"

    def __init__(self):
        self.launches = 0

    def build_csr(self, edge_index, n_nodes, validate=True):
        if validate and edge_index.numel() and (int(edge_index.min()) < 0 or int(edge_index.max()) >= n_nodes):
            raise ValueError(f"edge_index has edge(s) with a node id outside [0, {n_nodes})")
        row64 = edge_index[0]
        perm = torch.argsort(row64, stable=True)
        row = row64[perm].to(torch.int32)
        col = edge_index[1][perm].to(torch.int32)
        counts = torch.bincount(row64, minlength=n_nodes)
        rowptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=edge_index.device)
        rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        if validate == "defer":
            return rowptr, row, col, perm.to(torch.int32), None
        return rowptr, row, col, perm.to(torch.int32)

    def gather_rows(self, src, perm):
        return src[perm.long()].contiguous()

    def embed(self, dims, node_feat, node_loc, data_batch, emb_wt, emb_b, layer0, h, x4, batch32, P, Q,
              Hn, vsum, n_invalid=None):
        N, B, Fn, A, C, Na = dims
        if n_invalid is not None and N:
            bad = (data_batch < 0) | (data_batch >= B)
            bad[1:] |= data_batch[1:] < data_batch[:-1]
            n_invalid += int(bad.sum())
            data_batch = data_batch.clamp(0, B - 1)
        h.copy_(node_feat @ emb_wt + emb_b)
        x4.zero_()
        x4[:, :3] = node_loc
        batch32.copy_(data_batch.to(torch.int32))
        f = _fields(layer0, A, C, Na)
        P.copy_(h @ f["E_W1A"] + f["E_B1"])
        Q.copy_(h @ f["E_W1B"])
        Hn.copy_(h @ f["V_W1H"])
        vsum[:, 0:3].index_add_(0, data_batch, node_loc)
        vsum[:, 3].index_add_(0, data_batch, torch.ones(N, dtype=vsum.dtype, device=vsum.device))

    def edge_layer(self, dims, flags, row, col, ea, x4, P, Q, lp, agg_m, agg_x, n_edges_dev=None):
        N, E, A, C, Na = dims
        if E == 0:
            return
        f = _fields(lp, A, C, Na)
        r, c = row.long(), col.long()
        dx = x4[r, :3] - x4[c, :3]
        radial = (dx ** 2).sum(1, keepdim=True)
        if flags & _lib.FLAG_NORMALIZE:
            dx = dx / (radial.sqrt() + 1e-8)
        pre = P[r] + Q[c] + radial * f["E_W1R"]
        if A:
            pre = pre + ea @ f["E_W1E"]
        m = F.silu(F.silu(pre) @ f["E_W2"] + f["E_B2"])
        phi = F.silu(m @ f["E_WC"] + f["E_BC"]) @ f["E_W3"]
        if not flags & _lib.FLAG_LAST:
            agg_m.index_add_(0, r, m)
        agg_x[:, :3].index_add_(0, r, dx * phi.unsqueeze(1))

    def virtual_layer(self, dims, flags, batch32, x4, Hn, Xv, G, lp, agg_v, trans_v, vsum):
        N, B, A, C, Na = dims
        f = _fields(lp, A, C, Na)
        b = batch32.long()
        dX = Xv[b] - x4[:, :3].unsqueeze(-1)                       # [N,3,C]
        vr = dX.norm(dim=1)                                        # [N,C]
        pre = Hn.unsqueeze(1) + G[b] + vr.unsqueeze(-1) * f["V_W1R"]   # [N,C,64]
        mv = F.silu(F.silu(pre) @ f["V_W2"] + f["V_B2"])           # [N,C,64]
        phi_xv = F.silu(mv @ f["V_WXV"] + f["V_BXV"]) @ f["V_W3XV"]    # [N,C]
        phi_x = F.silu(mv @ f["V_WX"] + f["V_BX"]) @ f["V_W3X"]
        trans_v.zero_()
        trans_v[:, :3] = (-dX * phi_xv.unsqueeze(1)).mean(-1)
        vsum[:, 4:4 + 3 * C].index_add_(0, b, (dX * phi_x.unsqueeze(1)).reshape(N, 3 * C))
        if not flags & _lib.FLAG_LAST:
            agg_v.copy_(mv.mean(1))
            vsum[:, 4 + 3 * C:].index_add_(0, b, mv.reshape(N, C * H))

    def node_layer(self, dims, flags, rowptr, batch32, h, x4, vel, attr, agg_m, agg_x, agg_v, trans_v,
                   lp, lp_next, h_out, x4_out, P, Q, Hn, loc_out, vsum):
        N, B, A, C, Na = dims
        f = _fields(lp, A, C, Na)
        b = batch32.long()
        deg = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(h.dtype).unsqueeze(1)
        phiv = F.silu(h @ f["L_W"] + f["L_B"]) @ f["L_W3"] + f["L_B3"]
        xn = x4[:, :3] + agg_x[:, :3] / deg + trans_v[:, :3] + phiv.unsqueeze(1) * vel
        if not flags & _lib.FLAG_LAST:
            cat = [h, agg_m / deg, agg_v] + ([attr] if Na else [])
            hn = h + F.silu(torch.cat(cat, 1) @ f["N_W1"] + f["N_B1"]) @ f["N_W2"] + f["N_B2"]
            g = _fields(lp_next, A, C, Na)
            P.copy_(hn @ g["E_W1A"] + g["E_B1"])
            Q.copy_(hn @ g["E_W1B"])
            Hn.copy_(hn @ g["V_W1H"])
            h_out.copy_(hn)
        x4_out[:, :3] = xn
        if loc_out is not None:
            loc_out.copy_(xn)
        vsum[:, 0:3].index_add_(0, b, xn)
        vsum[:, 3].index_add_(0, b, torch.ones(N, dtype=vsum.dtype, device=vsum.device))
        if flags & _lib.FLAG_ZERO_AGG:
            agg_x.zero_()
            if agg_m is not None:
                agg_m.zero_()

    def allreduce_packed(self, comm, buf):
        comm.all_reduce(buf)

    def virtual_update(self, dims, flags, vsum, Xv, Hv, lp, lp_next, G, init_loc_mean=None, init_hv0=None, comm=None):
        B, A, C, Na = dims
        init, last = bool(flags & _lib.FLAG_INIT), bool(flags & _lib.FLAG_LAST)
        if comm is not None:
            comm.all_reduce(vsum)
        vsum_live = vsum
        vsum = vsum.clone()
        if flags & _lib.FLAG_ZERO_VSUM:
            vsum_live.zero_()
        if init and init_loc_mean is not None:
            Xv.copy_(init_loc_mean.unsqueeze(-1).expand(B, 3, C))
        if init and init_hv0 is not None and Hv is not None:
            Hv.copy_(init_hv0.unsqueeze(0).expand(B, C, H))
        n = vsum[:, 3].clamp(min=1)
        if not init:
            Xv += vsum[:, 4:4 + 3 * C].reshape(B, 3, C) / n.view(B, 1, 1)
        if last:
            return
        if not init:
            f = _fields(lp, A, C, Na)
            agg = vsum[:, 4 + 3 * C:].reshape(B, C, H) / n.view(B, 1, 1)
            Hv += F.silu(torch.cat([Hv, agg], -1) @ f["M_W1"] + f["M_B1"]) @ f["M_W2"] + f["M_B2"]
        g = _fields(lp_next, A, C, Na)
        xbar = vsum[:, 0:3] / n.view(B, 1)
        Z = Xv - xbar.unsqueeze(-1)                                # [B,3,C]
        mX = torch.einsum("bdi,bdj->bij", Z, Z)                    # [B,C,C]
        # G[b,c,:] = Hv[b,c,:]·W1v_V + Σ_j m_X[b,j,c]·W1v_M[j,:] + b1v
        G.copy_(Hv @ g["V_W1V"] + torch.einsum("bjc,jn->bcn", mX, g["V_W1M"]) + g["V_B1"])

    # ---- backward stand-ins (torch.autograd over tests/shadow_autograd.py): same accumulate / write contracts as
    # distegnn_edge_layer_bwd / distegnn_virtual_layer_bwd, so that FastEGNN's training path can run on CPU ---------
    def edge_layer_bwd(self, dims, flags, row, col, ea, x4, P, Q, lp, g_agg_m, g_agg_x, g_P, g_Q, g_x4, g_lp, n_edges_dev=None):
        from tests import shadow_autograd as sa
        N, E, A, C, Na = dims
        if E == 0:
            return
        with torch.enable_grad():
            Pl, Ql, xl, lpl = (t.detach().clone().requires_grad_(True) for t in (P, Q, x4[:, :3], lp))
            am, ax = sa.edge_stage(dims, flags, row, col, ea, xl, Pl, Ql, lpl)
            loss = (ax * g_agg_x[:, :3]).sum()
            if g_agg_m is not None and not flags & _lib.FLAG_LAST:
                loss = loss + (am * g_agg_m).sum()
            gP, gQ, gx, glp = torch.autograd.grad(loss, (Pl, Ql, xl, lpl), allow_unused=True)
        g_P += gP
        g_Q += gQ
        g_x4[:, :3] += gx
        g_lp += glp

    def virtual_bwd_prepare(self, A, Cn, Na, lp):
        return None

    def node_layer_bwd(self, dims, flags, rowptr, batch32, h, vel, attr, agg_m, agg_v, lp, lp_next, g_x_out, g_vsum,
                       g_h_out, g_P, g_Q, g_Hn, g_h, g_x, g_agg_x, g_trans_v, g_agg_m, g_agg_v, g_lp, g_lp_next):
        """torch.autograd through distegnn_b200._dense_stages.node_stage: the contract of distegnn_node_layer_bwd."""
        from tests import dense_stages as ds
        N, B, A, C, Na = dims
        last = bool(flags & _lib.FLAG_LAST)
        deg = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(h.dtype).unsqueeze(1)
        leaf = lambda t: t.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            hl, lpl = leaf(h), leaf(lp)
            xl, axl, tvl = (torch.zeros(N, 3, dtype=h.dtype, device=h.device).requires_grad_(True) for _ in range(3))
            aml, avl, lpn = (None, None, None) if last else (leaf(agg_m), leaf(agg_v), leaf(lp_next))
            xn, hn, Pn, Qn, Hnn = ds.node_stage(hl, xl, vel, attr if Na else None, aml, axl, avl, tvl, deg,
                                                ds.field_views(lpl, A, C, Na),
                                                None if last else ds.field_views(lpn, A, C, Na))
            gx = g_x_out if g_vsum is None else g_x_out + g_vsum[batch32.long(), 0:3]
            outs, gouts = [xn], [gx]
            for o, g in ((hn, g_h_out), (Pn, g_P), (Qn, g_Q), (Hnn, g_Hn)):
                if o is not None and g is not None:
                    outs.append(o)
                    gouts.append(g)
            ins = [hl, xl, axl, tvl, lpl] + ([] if last else [aml, avl, lpn])
            r = torch.autograd.grad(outs, ins, gouts, allow_unused=True)
            r = [torch.zeros_like(i) if g is None else g for g, i in zip(r, ins)]
        g_h.copy_(r[0])
        g_x.copy_(r[1])
        g_agg_x.zero_(); g_agg_x[:, :3] = r[2]
        g_trans_v.zero_(); g_trans_v[:, :3] = r[3]
        g_lp += r[4]
        if not last:
            g_agg_m.copy_(r[5])
            g_agg_v.copy_(r[6])
            g_lp_next += r[7]

    def virtual_update_bwd(self, dims, flags, vsum, Xv, Hv, lp, lp_next, g_Xn, g_Hn, g_G, g_vsum, g_Xv, g_Hv, g_lp, g_lp_next):
        """torch.autograd through _dense_stages.virtual_update_stage: the contract of distegnn_virtual_update_bwd."""
        from tests import dense_stages as ds
        B, A, C, Na = dims
        init, last = bool(flags & _lib.FLAG_INIT), bool(flags & _lib.FLAG_LAST)
        leaf = lambda t: t.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            vs, Xl = leaf(vsum), leaf(Xv)
            Hl = None if last else leaf(Hv)
            lpl = None if (init or lp is None) else leaf(lp)
            lpn = None if last else leaf(lp_next)
            Xn, Hn, Gn = ds.virtual_update_stage(vs, Xl, Hl, None if lpl is None else ds.field_views(lpl, A, C, Na),
                                                 None if last else ds.field_views(lpn, A, C, Na), init, C)
            pairs = [(o, g) for o, g in ((Xn, g_Xn), (Hn, g_Hn), (Gn, g_G)) if o is not None and g is not None and o.requires_grad]
            ins = [t for t in (vs, Xl, Hl, lpl, lpn) if t is not None]
            r = torch.autograd.grad([o for o, _ in pairs], ins, [g for _, g in pairs], allow_unused=True) if pairs else [None] * len(ins)
            r = {id(i): (torch.zeros_like(i) if g is None else g) for g, i in zip(r, ins)}
        g_vsum.copy_(r[id(vs)])
        g_Xv.copy_(r[id(Xl)])
        if Hl is not None and g_Hv is not None:
            g_Hv.copy_(r[id(Hl)])
        if lpl is not None:
            g_lp += r[id(lpl)]
        if lpn is not None:
            g_lp_next += r[id(lpn)]

    def embed_bwd(self, dims, node_feat, h0, lp0, g_h, g_P, g_Q, g_Hn, g_emb_wt, g_emb_b, g_lp0):
        from tests import dense_stages as ds
        N, B, Fn, A, C, Na = dims
        with torch.enable_grad():
            wl = torch.zeros(Fn, H, dtype=h0.dtype, device=h0.device).requires_grad_(True)
            bl = torch.zeros(H, dtype=h0.dtype, device=h0.device).requires_grad_(True)
            lp = lp0.detach().clone().requires_grad_(True)
            # h0 = feat·W + b is linear in (W, b): evaluate the stage around the saved h0
            hh = h0.detach() + node_feat @ wl + bl
            P, Q, Hn = ds.projections(hh, ds.field_views(lp, A, C, Na))
            r = torch.autograd.grad([hh, P, Q, Hn], [wl, bl, lp], [g_h, g_P, g_Q, g_Hn], allow_unused=True)
        g_emb_wt += r[0]
        g_emb_b += r[1]
        g_lp0 += r[2]

    def virtual_layer_bwd(self, dims, flags, batch32, x4, Hn, Xv, G, lp, wT, g_agg_v, g_trans_v, g_vsum, g_Hn, g_xv,
                          g_G, g_Xv, g_lp):
        from tests import shadow_autograd as sa
        N, B, A, C, Na = dims
        with torch.enable_grad():
            xl, Hl, Xl, Gl, lpl = (t.detach().clone().requires_grad_(True) for t in (x4[:, :3], Hn, Xv, G, lp))
            av, tv, tail = sa.virtual_stage(dims, flags, batch32, xl, Hl, Xl, Gl, lpl)
            loss = (tv * g_trans_v[:, :3]).sum() + (tail[:, :3 * C] * g_vsum[:, 4:4 + 3 * C]).sum()
            if g_agg_v is not None and not flags & _lib.FLAG_LAST:
                loss = loss + (av * g_agg_v).sum() + (tail[:, 3 * C:] * g_vsum[:, 4 + 3 * C:]).sum()
            gx, gH, gX, gG, glp = torch.autograd.grad(loss, (xl, Hl, Xl, Gl, lpl), allow_unused=True)
        g_Hn.copy_(gH)
        g_xv.zero_()
        g_xv[:, :3] = gx
        g_G += gG
        g_Xv += gX
        g_lp += glp
