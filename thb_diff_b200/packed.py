"""PackedHierarchy -- the flat, device-resident form of a THB space that the CUDA kernels consume.

The reference keeps the hierarchy as Python dicts: per-cell support lists
(``compute_active_cells_active_supp``, THB/core.py:14-40,424-449) and dense per-level coefficient tables of
shape [*sh_fns[lev], *sh_fns[max_lev]] (``compute_refinement_operators``, THB/core.py:43-114; 3.9 GB for
the README 3-D toy, ~1e11 entries for BASELINE config 3).  Here the same information is stored as

  knots      f32  all levels / dims concatenated                     (staged in shared memory by kernels)
  cell_slot  i32  per level, one entry per cell: slot id of the active cell or -1
  cellinfo   i32  [nslots, 8] = level, c0, c1, c2, supp_off, S, tile_off, n_direct
  dmask      u64  [nslots]  which of the cell's own-level (p+1)^d window functions are active
  supp_jm    i32  CSR of flat support ids (reference order: own level first, then ancestors, C order)
  tiles      f32  [ntiles, TS]  for every (active cell, coarser-level support): the (p+1)^d coefficients
                  of that (single-level-truncated) function in the CELL's-level basis on the cell

PHI(point, support) = <tile, tensor product of the cell-level B-splines at the point>.  In exact
arithmetic this equals the reference's finest-level formulation (the refinement relation is exact); the
own-level supports have one-hot tiles, so they are not stored -- their PHI is the tensor-product entry
itself (``onehot_direct=False`` stores them explicitly, which reproduces the reference's dense
gather-multiply operation count).

Semantics follow the reference: single-level truncation (core.py:98-111); the finest level
uses the identity (the correct reading of core.py:113); no fp16 quantisation.
"""
import os

import numpy as np
import torch

MAX_LEVELS = 8
INFO_W = 12  # ints per cellinfo row


def _ceil4(x):
    return (x + 3) // 4 * 4


def _span_poly_tables(kv, degrees, nlevels):
    """Per level, dimension and 1-D cell: the p+1 B-splines that live on the cell as polynomials in the cell-local coordinate
    s = (u - lo) / half - 1 in [-1, 1]:  N_r(u) = sum_m C[m][r] s^m  (float64 Cox-de Boor at Chebyshev nodes + an exact
    (p+1)-point interpolation, rounded to float32 once).  Rows of 20 floats: C as [m][r] padded to 4 x 4, then
    (lo, 1 / half, 0, 0) with lo the cell's left breakpoint and half its half width.  Lets the kernels evaluate a cell's geometry in MONOMIAL form (nested Horner, no per-point
    Cox-de Boor); the centred, scaled variable keeps the coefficients O(1): float32 Horner stays ~1e-6 of max|geometry|
    (measured 1.6e-6 worst over random nets)."""
    d = len(degrees)
    if any(int(p) > 3 for p in degrees):
        return None, None   # rows hold 4 x 4 coefficients: degrees above three have no monomial tables (and no fused kernels)
    rows, off = [], np.zeros((nlevels, 3), dtype=np.int64)
    pos = 0
    for l in range(nlevels):
        for k in range(d):
            p = int(degrees[k])
            U = kv[l][k].astype(np.float32).astype(np.float64)   # the knots the device sees
            nc = len(U) - 2 * p - 1
            lo, hi = U[p:p + nc], U[p + 1:p + nc + 1]
            cen, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
            nodes = np.cos((2 * np.arange(p + 1) + 1) * np.pi / (2 * (p + 1)))            # Chebyshev nodes in (-1, 1)
            x = cen[:, None] + half[:, None] * nodes[None, :]                             # [nc, p+1]
            # Cox-de Boor (A2.2) at span = cell + p, vectorised over cells and nodes: N[..., r], r = 0..p
            N = np.zeros(x.shape + (p + 1,))
            N[..., 0] = 1.0
            span = np.arange(nc) + p
            for j in range(1, p + 1):
                saved = np.zeros(x.shape)
                for r in range(j):
                    right = U[span + r + 1][:, None] - x
                    left = x - U[span + 1 - j + r][:, None]
                    t = N[..., r] / (right + left)
                    N[..., r] = saved + right * t
                    saved = left * t
                N[..., j] = saved
            Vinv = np.linalg.inv(np.vander(nodes, p + 1, increasing=True))               # [m][node]
            C = np.einsum("mq,cqr->cmr", Vinv, N)                                         # [cell][m][r]
            row = np.zeros((nc, 20))
            blk = np.zeros((nc, 4, 4))
            blk[:, :p + 1, :p + 1] = C
            row[:, :16] = blk.reshape(nc, 16)
            row[:, 16], row[:, 17] = lo, 1.0 / half   # lo is a float32 knot: u - lo is formed without an absolute rounding error
            rows.append(row.astype(np.float32))
            off[l, k] = pos
            pos += nc
    return np.concatenate(rows), off


class PackedHierarchy:
    def __init__(self):
        pass

    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_space(cls, space, device="cuda", onehot_direct=True, chunk=1 << 18, truncation="single"):
        if space.fns is None:
            raise ValueError("call Space.build_hierarchy_from_domain_sequence() first")
        return cls.build(space.knotvectors, space.degrees, space.cells, space.fns, space.Coeff,
                         device=device, onehot_direct=onehot_direct, chunk=chunk, truncation=truncation)

    @classmethod
    def build(cls, knotvectors, degrees, cells, fns, Coeff=None, device="cuda", onehot_direct=True, chunk=1 << 18,
              build_device=None, truncation="single"):
        self = cls()
        self.version = 0
        L = max(knotvectors.keys()) + 1
        d = len(degrees)
        if d not in (2, 3):
            raise ValueError("only 2-D and 3-D tensor products are supported")
        if L > MAX_LEVELS:
            raise ValueError(f"at most {MAX_LEVELS} levels")
        p = tuple(int(x) for x in degrees)
        T = int(np.prod([x + 1 for x in p]))
        if T > 64:
            raise ValueError("tensor-product size (p+1)^d must be <= 64")
        TS = _ceil4(T)
        self.ndim, self.nlevels, self.degrees, self.T, self.TS = d, L, p, T, TS
        self.onehot_direct = bool(onehot_direct)
        if truncation not in ("single", "recursive"):
            raise ValueError("truncation must be 'single' (the reference's, THB/core.py:98-111) or 'recursive'")
        self.truncation = truncation
        bdev = torch.device(build_device) if build_device is not None else torch.device(device)

        kv = {l: [np.asarray(U, dtype=np.float64) for U in knotvectors[l]] for l in range(L)}
        sh_cells, sh_fns = {}, {}
        for l in range(L):
            for k in range(d):
                U = kv[l][k]
                if len(np.unique(U)) != len(U) - 2 * p[k]:
                    raise ValueError("knot vectors must be open with simple interior knots")
                if np.any(U.astype(np.float32)[p[k]:len(U) - p[k] - 1] >= U.astype(np.float32)[p[k] + 1:len(U) - p[k]]):
                    raise ValueError("knot spans collapse in float32")
            sh_cells[l] = tuple(len(kv[l][k]) - 2 * p[k] - 1 for k in range(d))
            sh_fns[l] = tuple(c + q for c, q in zip(sh_cells[l], p))
            if tuple(cells[l].shape) != sh_cells[l] or tuple(fns[l].shape) != sh_fns[l]:
                raise ValueError(f"mask shapes at level {l} do not match the knot vectors")
            if l > 0:
                for k in range(d):
                    if sh_cells[l][k] != 2 * sh_cells[l - 1][k] or not np.array_equal(
                            np.unique(kv[l][k])[::2], np.unique(kv[l - 1][k])):
                        raise ValueError("levels must be dyadically nested (refine_knotvector)")
        self.sh_cells, self.sh_fns = sh_cells, sh_fns
        fn_off = np.zeros(L + 1, dtype=np.int64)
        for l in range(L):
            fn_off[l + 1] = fn_off[l] + int(np.prod(sh_fns[l]))
        self.fn_off = fn_off

        # ---- knots
        knot_off = np.zeros((MAX_LEVELS, 3), dtype=np.int32)
        knot_len = np.zeros((MAX_LEVELS, 3), dtype=np.int32)
        flat = []
        pos = 0
        for l in range(L):
            for k in range(d):
                knot_off[l, k], knot_len[l, k] = pos, len(kv[l][k])
                flat.append(kv[l][k].astype(np.float32))
                pos += len(kv[l][k])
        self.knot_off, self.knot_len = knot_off, knot_len
        knots = np.concatenate(flat)

        # ---- active cells -> slots
        act_idx = {}
        slot_base = np.zeros(L + 1, dtype=np.int64)
        cs_off = np.zeros(MAX_LEVELS, dtype=np.int64)
        tot = 0
        for l in range(L):
            cs_off[l] = tot
            tot += int(np.prod(sh_cells[l]))
            act_idx[l] = np.argwhere(np.asarray(cells[l]) == 1).astype(np.int64)
            slot_base[l + 1] = slot_base[l] + len(act_idx[l])
        nslots = int(slot_base[L])
        if nslots == 0:
            raise ValueError("no active cells")
        cell_slot = np.full(tot, -1, dtype=np.int32)
        for l in range(L):
            if len(act_idx[l]):
                lin = np.ravel_multi_index(tuple(act_idx[l].T), sh_cells[l])
                cell_slot[cs_off[l] + lin] = slot_base[l] + np.arange(len(lin))
        self.cell_slot_off, self.nslots = cs_off, nslots
        # span lookup tables of the finest level: uniform bins at most 0.4 cells wide; entry = cell of the bin's left
        # edge, so the true cell of a point is that cell or the next one (the kernel compares once and verifies)
        lut_parts, lut_off, lut_n, lut_scale = [], [0, 0, 0], [0, 0, 0], [0.0, 0.0, 0.0]
        lut_exact_parts = []
        lpos = 0
        for k in range(d):
            B32 = kv[L - 1][k].astype(np.float32)[p[k]:len(kv[L - 1][k]) - p[k]]
            Bk = B32.astype(np.float64)
            nb = int(np.ceil(2.5 * (Bk[-1] - Bk[0]) / np.min(np.diff(Bk))))
            if nb > 8192:
                lut_parts = None
                break
            sc = np.float32(nb / (Bk[-1] - Bk[0]))
            left = Bk[0] + (np.arange(nb) - 0.05) / float(sc)
            lut_parts.append(np.clip(np.searchsorted(Bk, left, side="right") - 1, 0, len(Bk) - 2).astype(np.int32))
            # PROVABLE form.  bin(x) = min(int((x - u0) * sc), nb - 1) in float32 is monotone in x, so with
            # bin_c = bin(B[c]) of the interior breakpoints:  bin(x) > bin_c => x >= B[c],  bin(x) < bin_c => x < B[c].
            # If no two breakpoints share a bin, the cell of every x of bin b is lut[b] = #{c : bin_c < b} or the next one,
            # and one comparison with B[lut[b] + 1] decides -- no verification needed on the device.
            if lut_exact_parts is not None and len(B32) >= 2:
                with np.errstate(all="ignore"):
                    bin_c = np.minimum(((B32[1:-1] - B32[0]) * sc).astype(np.int32), nb - 1)   # float32 arithmetic, as on the device
                if len(bin_c) == 0 or (np.all(np.diff(bin_c) > 0) and bin_c[0] >= 0):
                    lut_exact_parts.append(np.searchsorted(bin_c, np.arange(nb), side="left").astype(np.int32))
                else:
                    lut_exact_parts = None
            lut_off[k], lut_n[k], lut_scale[k] = lpos, nb, float(sc)
            lpos += nb
        self.lut_exact = bool(lut_parts) and lut_exact_parts is not None and os.environ.get("THB_LUT_EXACT", "1") != "0"
        if self.lut_exact:
            lut_parts = lut_exact_parts   # also valid for the verifying kernels: the cell is lut[bin] or lut[bin] + 1
        self._lut = (np.concatenate(lut_parts), lut_off, lut_n, lut_scale) if lut_parts else None
        # finest-cell -> active-slot map (every finest cell lies in exactly one active cell): lets the kernels
        # find a point's cell with ONE lookup instead of one per level.  Skipped for very deep hierarchies.
        finest_map = None
        if int(np.prod(sh_cells[L - 1])) <= (1 << 26):
            finest_map = np.full(sh_cells[L - 1], -1, dtype=np.int32)
            for l in range(L):
                if not len(act_idx[l]):
                    continue
                grid = cell_slot[cs_off[l]:cs_off[l] + int(np.prod(sh_cells[l]))].reshape(sh_cells[l])
                f = 1 << (L - 1 - l)
                up = grid
                for k in range(d):
                    up = np.repeat(up, f, axis=k)
                np.maximum(finest_map, up, out=finest_map)

        # ---- admissibility: no active level-(l+1) function may overlap an active level-l cell
        fmask = {l: (np.asarray(fns[l]) == 1) for l in range(L)}
        for l in range(L - 1):
            c = act_idx[l]
            if not len(c):
                continue
            nxt = fmask[l + 1]
            hit = np.zeros(len(c), dtype=bool)
            rngs = [range(p[k] + 2) for k in range(d)]
            for off in np.array(np.meshgrid(*rngs, indexing="ij")).reshape(d, -1).T:
                hit |= nxt[tuple((2 * c[:, k] + off[k]) for k in range(d))]
            if hit.any():
                raise ValueError(f"inadmissible hierarchy: an active level-{l + 1} function overlaps an active level-{l} cell")

        # ---- window offsets in C order
        wt = np.array(np.meshgrid(*[range(q + 1) for q in p], indexing="ij")).reshape(d, -1).T.astype(np.int64)  # [T,d]

        info = np.zeros((nslots, INFO_W), dtype=np.int32)
        dmask = np.zeros(nslots, dtype=np.uint64)
        jm_parts, batches = [], []
        supp_pos, tile_pos = 0, 0
        for lev in range(L):
            c = act_idx[lev]
            n = len(c)
            if n == 0:
                continue
            flags, jms, fidx = [], [], []
            for l in range(lev, -1, -1):
                a = c >> (lev - l)
                w = a[:, None, :] + wt[None, :, :]                                # [n,T,d]
                lin = np.ravel_multi_index(tuple(w[..., k] for k in range(d)), sh_fns[l])
                flags.append(fmask[l].reshape(-1)[lin])
                jms.append(fn_off[l] + lin)
                fidx.append(w)
            F = np.concatenate(flags, axis=1)
            J = np.concatenate(jms, axis=1)
            S = F.sum(axis=1)
            own = flags[0]
            ndir = own.sum(axis=1) if onehot_direct else np.zeros(n, dtype=np.int64)
            bits = (own.astype(np.uint64) << np.arange(T, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)
            sl = slice(int(slot_base[lev]), int(slot_base[lev + 1]))
            info[sl, 0] = lev
            info[sl, 1:1 + d] = c
            soff = supp_pos + np.concatenate([[0], np.cumsum(S)[:-1]])
            info[sl, 4], info[sl, 5], info[sl, 7] = soff, S, ndir
            dmask[sl] = bits
            jm_parts.append(J[F])
            supp_pos += int(S.sum())
            # tiled supports, reference order
            first = 1 if onehot_direct else 0
            cnts = np.stack([flags[j].sum(axis=1) for j in range(first, lev + 1)], axis=1) if lev + 1 > first else np.zeros((n, 0), dtype=np.int64)
            ntile = cnts.sum(axis=1)
            toff = tile_pos + np.concatenate([[0], np.cumsum(ntile)[:-1]])
            info[sl, 6] = toff
            base = toff[:, None] + np.concatenate([np.zeros((n, 1), dtype=np.int64), np.cumsum(cnts, axis=1)[:, :-1]], axis=1) if cnts.shape[1] else None
            for j in range(first, lev + 1):
                l = lev - j
                fl = flags[j]
                rows, cols = np.nonzero(fl)
                if len(rows) == 0:
                    continue
                rank = (np.cumsum(fl, axis=1) - 1)[rows, cols]
                gidx = base[rows, j - first] + rank
                batches.append((lev, l, c[rows], fidx[j][rows, cols], gidx))
            tile_pos += int(ntile.sum())
        ntiles = tile_pos
        supp_jm = np.concatenate(jm_parts).astype(np.int32)
        self.max_S = int(info[:, 5].max())
        self.max_tiled = int((info[:, 5] - info[:, 7]).max())
        self.ntiles, self.nsupp = ntiles, len(supp_jm)

        self._batches, self._fmask, self._kv_shapes = batches, fmask, None
        self._chunk, self._bdev = chunk, bdev
        self.tiles = None
        dev = torch.device(device)
        self.device = dev
        self.cellinfo_host = info
        self.knots = torch.from_numpy(knots).to(dev)
        poly, self.poly_off = _span_poly_tables(kv, p, L)
        self.poly = torch.from_numpy(poly).to(dev) if poly is not None else None
        self.cell_slot = torch.from_numpy(cell_slot).to(dev)
        self.finest_slot = torch.from_numpy(finest_map.reshape(-1)).to(dev) if finest_map is not None else None
        self.span_lut = torch.from_numpy(self._lut[0]).to(dev) if self._lut is not None else None
        self.cellinfo = torch.from_numpy(info).to(dev)
        self.dmask = torch.from_numpy(dmask.view(np.int64)).to(dev)
        self.supp_jm = torch.from_numpy(supp_jm).to(dev)
        self.tiles = torch.zeros((1, TS), dtype=torch.float32, device=dev)
        self.has_tiles = ntiles == 0
        if Coeff is not None:
            self.attach_tiles(Coeff)
        self.nCP = int(fn_off[L])
        return self


    def _recursive_tiles(self, lev, l, c_, f_, Ct, fmask, ar, bdev):
        """Tiles under the PROPER (recursive) truncation: the level-l function is written in level l+1, the active
        level-(l+1) functions are removed, the rest is refined to level l+2, the active level-(l+2) functions are
        removed, ... up to the cell's level.  Everything stays inside the (p+1)^d window of the cell's ancestors:
        a function that is non-zero on the cell only has parents that are non-zero on it."""
        d, p, T, TS = self.ndim, self.degrees, self.T, self.TS
        m = len(c_)
        let = "abc"[:d]

        def window_mask(level, a):
            idx = [a[:, k, None] + ar[k][None, :] for k in range(d)]
            fm = torch.as_tensor(fmask[level], device=bdev)
            if d == 2:
                return fm[idx[0][:, :, None], idx[1][:, None, :]]
            return fm[idx[0][:, :, None, None], idx[1][:, None, :, None], idx[2][:, None, None, :]]

        # level l+1: subdivision weights of function f in the window of the cell's level-(l+1) ancestor
        a = c_ >> (lev - (l + 1))
        W = [Ct[(l, k)][f_[:, k, None], a[:, k, None] + ar[k][None, :]] for k in range(d)]
        w = torch.einsum(",".join("m" + x for x in let) + "->m" + let, *W)
        w = w * (~window_mask(l + 1, a)).to(w.dtype)
        for lv in range(l + 2, lev + 1):
            a_prev, a = a, c_ >> (lev - lv)
            # local blocks of the 1-D subdivision matrices: rows = previous window, columns = this window
            blocks = [Ct[(lv - 1, k)][(a_prev[:, k, None] + ar[k][None, :])[:, :, None], (a[:, k, None] + ar[k][None, :])[:, None, :]] for k in range(d)]
            up = "xyz"[:d]
            w = torch.einsum("m" + let + "," + ",".join("m" + x + y for x, y in zip(let, up)) + "->m" + up, w, *blocks)
            w = w * (~window_mask(lv, a)).to(w.dtype)
        t = torch.zeros((m, TS), dtype=torch.float32, device=bdev)
        t[:, :T] = w.reshape(m, T).to(torch.float32)
        return t, w

    def _rank_one(self, tile):
        """tile [m, n0, n1(, n2)] float64 -> (is rank one [m] bool, factor vectors [m, 12] f32: 4 floats per dimension,
        zero padded; the third vector is (1,0,0,0) in 2-D).  Rank one = the tile is the outer product of the fibres
        through its largest entry (exact for untruncated functions and for truncation across a flat face)."""
        d, p = self.ndim, self.degrees
        m = tile.shape[0]
        dev = tile.device
        flat = tile.reshape(m, -1)
        idx = flat.abs().argmax(dim=1)
        piv = flat[torch.arange(m, device=dev), idx]
        n = [q + 1 for q in p]
        if d == 2:
            i0, j0 = idx // n[1], idx % n[1]
            a = tile[torch.arange(m, device=dev), :, j0]
            b = tile[torch.arange(m, device=dev), i0, :]
            safe = torch.where(piv == 0, torch.ones_like(piv), piv)
            b = b / safe[:, None]
            rec = torch.einsum("ma,mb->mab", a, b)
            vecs = [a, b, None]
        else:
            i0, r = idx // (n[1] * n[2]), idx % (n[1] * n[2])
            j0, k0 = r // n[2], r % n[2]
            ar_ = torch.arange(m, device=dev)
            a = tile[ar_, :, j0, k0]
            b = tile[ar_, i0, :, k0]
            c = tile[ar_, i0, j0, :]
            safe = torch.where(piv == 0, torch.ones_like(piv), piv)
            b = b / safe[:, None]
            c = c / safe[:, None]
            rec = torch.einsum("ma,mb,mc->mabc", a, b, c)
            vecs = [a, b, c]
        ok = (rec - tile).reshape(m, -1).abs().max(dim=1).values <= 1e-12
        out = torch.zeros((m, 12), dtype=torch.float32, device=dev)
        for k in range(3):
            if vecs[k] is None:
                out[:, 4 * k] = 1.0
            else:
                out[:, 4 * k:4 * k + n[k]] = vecs[k].to(torch.float32)
        return ok, out

    def _split_tiled_supports(self, sepflag, sepvec):
        """Per cell: the rank-one tiled supports (12 floats each) and the others (full tiles), each list in the
        reference's relative order; offsets go into cellinfo columns 8..10."""
        info = self.cellinfo_host
        S, ndir, soff = info[:, 5].astype(np.int64), info[:, 7].astype(np.int64), info[:, 4].astype(np.int64)
        nt = S - ndir
        supp = self.supp_jm.cpu().numpy()
        pos_in_cell = np.arange(len(supp)) - np.repeat(soff, S)
        tiled = pos_in_cell >= np.repeat(ndir, S)
        jm_tiled = supp[tiled]                                   # aligned with the global tile order
        assert len(jm_tiled) == self.ntiles == len(sepflag)
        cell_of_tile = np.repeat(np.arange(len(info)), nt)
        nsep = np.bincount(cell_of_tile[sepflag], minlength=len(info)).astype(np.int64)
        nfull = nt - nsep
        info[:, 8] = np.concatenate([[0], np.cumsum(nsep)[:-1]])
        info[:, 9] = nsep
        info[:, 10] = np.concatenate([[0], np.cumsum(nfull)[:-1]])
        dev = self.device
        self.sep_vec = torch.from_numpy(np.ascontiguousarray(sepvec[sepflag])).to(dev) if sepflag.any() else torch.zeros((1, 12), dtype=torch.float32, device=dev)
        self.sep_jm = torch.from_numpy(np.ascontiguousarray(jm_tiled[sepflag]).astype(np.int32)).to(dev) if sepflag.any() else torch.zeros(1, dtype=torch.int32, device=dev)
        full = ~sepflag
        # where each tiled support (global tile order = reference order inside every cell) lives after the split
        ref = np.where(sepflag, (np.cumsum(sepflag) - 1) << 1, ((np.cumsum(full) - 1) << 1) | 1).astype(np.int32)
        self.tiled_ref = torch.from_numpy(ref).to(dev) if len(ref) else torch.zeros(1, dtype=torch.int32, device=dev)
        self.full_tiles = self.tiles[torch.from_numpy(np.nonzero(full)[0]).to(dev)].contiguous() if full.any() else torch.zeros((1, self.TS), dtype=torch.float32, device=dev)
        self.full_jm = torch.from_numpy(np.ascontiguousarray(jm_tiled[full]).astype(np.int32)).to(dev) if full.any() else torch.zeros(1, dtype=torch.int32, device=dev)
        self.cellinfo = torch.from_numpy(info).to(dev)
        self.nsep_total, self.nfull_total = int(nsep.sum()), int(nfull.sum())
        self.max_sep = int(nsep.max()) if len(nsep) else 0

    def attach_tiles(self, Coeff):
        """Computes the coefficient tiles of the coarser-level supports from the 1-D subdivision matrices
        Coeff[lev][dim] ([nfns(lev), nfns(lev+1)], reference multilevel_spline_space.py:78-92)."""
        if self.has_tiles and self.ntiles == 0:
            return self
        L, d, p, T, TS = self.nlevels, self.ndim, self.degrees, self.T, self.TS
        sh_fns, fmask, batches, chunk, bdev, ntiles = self.sh_fns, self._fmask, self._batches, self._chunk, self._bdev, self.ntiles
        # ---- 1-D refinement chains R[l][lev] = Coeff[l] ... Coeff[lev-1]
        R = {}
        for k in range(d):
            for l in range(L):
                R[(k, l, l)] = np.eye(sh_fns[l][k])
                for lev in range(l + 1, L):
                    R[(k, l, lev)] = R[(k, l, lev - 1)] @ np.asarray(Coeff[lev - 1][k], dtype=np.float64)

        tiles = torch.zeros((max(ntiles, 1), TS), dtype=torch.float32, device=bdev)
        sepflag = torch.zeros(max(ntiles, 1), dtype=torch.bool, device=bdev)
        sepvec = torch.zeros((max(ntiles, 1), 12), dtype=torch.float32, device=bdev)
        es_u = {2: "ma,mb->mab", 3: "ma,mb,mc->mabc"}[d]
        es_t = {2: "mab,max,mby->mxy", 3: "mabc,max,mby,mcz->mxyz"}[d]
        tdev = lambda x, dt=torch.float64: torch.as_tensor(x, device=bdev, dtype=dt)
        Rt = {key: tdev(v) for key, v in R.items()}
        Ct = {(l, k): tdev(np.asarray(Coeff[l][k], dtype=np.float64)) for l in range(L - 1) for k in range(d)}
        ar = [torch.arange(q + 1, device=bdev) for q in p]
        for lev, l, cc, ff, gidx in batches:
            for s in range(0, len(gidx), chunk):
                c_ = torch.as_tensor(cc[s:s + chunk], device=bdev)
                f_ = torch.as_tensor(ff[s:s + chunk], device=bdev)
                g_ = torch.as_tensor(gidx[s:s + chunk], device=bdev)
                if l == lev:
                    t = torch.zeros((len(g_), TS), dtype=torch.float32, device=bdev)
                    loc = f_ - c_
                    lin = loc[:, 0]
                    for k in range(1, d):
                        lin = lin * (p[k] + 1) + loc[:, k]
                    t[torch.arange(len(g_), device=bdev), lin] = 1.0
                    tiles[g_] = t
                    continue
                if self.truncation == "recursive":
                    t, w64 = self._recursive_tiles(lev, l, c_, f_, Ct, fmask, ar, bdev)
                    tiles[g_] = t
                    fl, vec = self._rank_one(w64)
                    sepflag[g_] = fl
                    sepvec[g_] = vec
                    continue
                A = [Rt[(k, l, lev)][f_[:, k, None], c_[:, k, None] + ar[k][None, :]] for k in range(d)]
                tile = torch.einsum(es_u, *A)
                e = l + 1
                a_ = c_ >> (lev - e)
                lin = a_[:, 0, None] + ar[0][None, :]
                widx = [a_[:, k, None] + ar[k][None, :] for k in range(d)]          # [m, p_k+1]
                if d == 2:
                    M = torch.as_tensor(fmask[e], device=bdev)[widx[0][:, :, None], widx[1][:, None, :]]
                else:
                    M = torch.as_tensor(fmask[e], device=bdev)[widx[0][:, :, None, None], widx[1][:, None, :, None], widx[2][:, None, None, :]]
                if bool(M.any()):
                    G = []
                    for k in range(d):
                        W = Ct[(l, k)][f_[:, k, None], widx[k]]                                        # [m, δ]
                        Rk = Rt[(k, e, lev)][widx[k][:, :, None], (c_[:, k, None] + ar[k][None, :])[:, None, :]]  # [m, δ, t]
                        G.append(W[:, :, None] * Rk)
                    tile = tile - torch.einsum(es_t, M.to(torch.float64), *G)
                t = torch.zeros((len(g_), TS), dtype=torch.float32, device=bdev)
                t[:, :T] = tile.reshape(len(g_), T).to(torch.float32)
                tiles[g_] = t
                fl, vec = self._rank_one(tile)
                sepflag[g_] = fl
                sepvec[g_] = vec

        self.version = getattr(self, "version", 0) + 1  # plans copy per-cell table offsets: older plans must be rebuilt
        self.tiles = tiles.to(self.device)
        self._split_tiled_supports(sepflag[:ntiles].cpu().numpy(), sepvec[:ntiles].cpu().numpy())
        self.has_tiles = True
        self._desc_cache = None
        return self

    # ------------------------------------------------------------------------------------------
    def active_rows(self):
        """(active_ids int32 [n_active], row_map int32 [nCP]): the flat ids of the functions that are a support of some
        active cell -- the only rows of ctrl_pts the evaluation reads and the only ones that ever receive a gradient --
        in ascending order, and the inverse map (-1 elsewhere).  Cached; device tensors."""
        c = getattr(self, "_active_rows", None)
        if c is None:
            ids = torch.unique(self.supp_jm.to(torch.int64))
            row_map = torch.full((self.nCP,), -1, dtype=torch.int32, device=self.device)
            row_map[ids] = torch.arange(ids.shape[0], dtype=torch.int32, device=self.device)
            c = self._active_rows = (ids.to(torch.int32), row_map)
        return c

    def ordered_scatter_tables(self):
        """(tr_ptr int32 [n_active + 1], tr_idx int32 [nsupp]) for the ORDERED backward (thb_eval_fused_backward_ordered):
        every (cell, support) pair owns one slot of a contribution buffer, slot = supp_off(cell) + position with the cell's
        supports in the order  own-level (window order) | rank-one tiled | full-tile;  tr_idx lists, for every active
        function (compact row of active_rows()), its slots in increasing slot number.  Cached; device tensors."""
        c = getattr(self, "_ordered_tables", None)
        if c is None:
            info = self.cellinfo.long()
            soff, S, ndir = info[:, 4], info[:, 5], info[:, 7]
            nsupp = int(self.supp_jm.shape[0])
            fn = torch.empty(nsupp, dtype=torch.int64, device=self.device)
            pos = torch.arange(nsupp, device=self.device)
            cell = torch.repeat_interleave(torch.arange(self.nslots, device=self.device), S)   # slots are laid out cell by cell
            local = pos - soff[cell]
            direct = local < ndir[cell]
            fn[direct] = self.supp_jm.long()[pos[direct]]
            if getattr(self, "sep_vec", None) is not None:
                sep_off, nsep, full_off = info[:, 8], info[:, 9], info[:, 10]
                t = local - ndir[cell]
                is_sep = (~direct) & (t < nsep[cell])
                is_full = (~direct) & (~is_sep)
                fn[is_sep] = self.sep_jm.long()[(sep_off[cell] + t)[is_sep]]
                fn[is_full] = self.full_jm.long()[(full_off[cell] + t - nsep[cell])[is_full]]
            else:
                fn[~direct] = self.supp_jm.long()[pos[~direct]]
            ids, row_map = self.active_rows()
            rows = row_map.long()[fn]
            order = torch.argsort(rows, stable=True)
            cnt = torch.bincount(rows, minlength=ids.shape[0])
            tr_ptr = torch.zeros(ids.shape[0] + 1, dtype=torch.int64, device=self.device)
            tr_ptr[1:] = torch.cumsum(cnt, 0)
            c = self._ordered_tables = (tr_ptr.to(torch.int32), order.to(torch.int32))
        return c

    @property
    def n_active(self):
        return int(self.active_rows()[0].shape[0])

    def nbytes(self):
        ts = [self.knots, self.cell_slot, self.cellinfo, self.dmask, self.supp_jm, self.tiles] + [t for t in (getattr(self, 'sep_vec', None), getattr(self, 'full_tiles', None)) if t is not None] + ([self.finest_slot] if self.finest_slot is not None else [])
        return sum(t.numel() * t.element_size() for t in ts)

    def summary(self):
        return dict(ndim=self.ndim, levels=self.nlevels, degrees=self.degrees, nslots=self.nslots, nCP=self.nCP,
                    nsupp=self.nsupp, ntiles=self.ntiles, max_S=self.max_S, max_tiled=self.max_tiled,
                    bytes=self.nbytes(), onehot_direct=self.onehot_direct)


def _desc(self):
    """ctypes thb_space_desc for this hierarchy (cached; keeps the tensors alive through self)."""
    d = getattr(self, "_desc_cache", None)
    if d is not None:
        return d
    from ._lib import SpaceDesc

    d = SpaceDesc()
    d.ndim, d.nlevels = self.ndim, self.nlevels
    for k in range(3):
        d.degree[k] = self.degrees[k] if k < self.ndim else 0
    d.tile_stride, d.nslots, d.max_tiled = self.TS, self.nslots, self.max_tiled
    for l in range(self.nlevels):
        for k in range(self.ndim):
            d.knot_off[l][k] = int(self.knot_off[l, k])
            d.knot_len[l][k] = int(self.knot_len[l, k])
            d.ncells[l][k] = int(self.sh_cells[l][k])
            d.nfns[l][k] = int(self.sh_fns[l][k])
        if self.ndim == 2:
            d.ncells[l][2], d.nfns[l][2] = 1, 1
        d.cell_slot_off[l] = int(self.cell_slot_off[l])
    for l in range(self.nlevels + 1):
        d.fn_off[l] = int(self.fn_off[l])
    d.knots, d.cell_slot, d.cellinfo = self.knots.data_ptr(), self.cell_slot.data_ptr(), self.cellinfo.data_ptr()
    d.dmask, d.supp_jm, d.tiles = self.dmask.data_ptr(), self.supp_jm.data_ptr(), self.tiles.data_ptr()
    d.finest_slot = self.finest_slot.data_ptr() if self.finest_slot is not None else None
    if getattr(self, "sep_vec", None) is not None:
        d.sep_vec, d.sep_jm = self.sep_vec.data_ptr(), self.sep_jm.data_ptr()
        d.full_tiles, d.full_jm = self.full_tiles.data_ptr(), self.full_jm.data_ptr()
    d.tiled_ref = self.tiled_ref.data_ptr() if getattr(self, "tiled_ref", None) is not None and getattr(self, "sep_vec", None) is not None else None
    d.poly = self.poly.data_ptr() if self.poly is not None else None
    if self.poly is not None:
        for l in range(self.nlevels):
            for k in range(self.ndim):
                d.poly_off[l][k] = int(self.poly_off[l, k])
    d.span_lut = self.span_lut.data_ptr() if self.span_lut is not None else None
    d.lut_exact = 1 if (self.span_lut is not None and getattr(self, "lut_exact", False)) else 0
    if self.span_lut is not None:
        for k in range(self.ndim):
            d.lut_off[k], d.lut_n[k], d.lut_scale[k] = self._lut[1][k], self._lut[2][k], self._lut[3][k]
    self._desc_cache = d
    return d


PackedHierarchy.desc = _desc
