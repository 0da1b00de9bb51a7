"""numpy interpreter of a dumped engine plan (qb_plan_dump).  TEST INFRASTRUCTURE ONLY.

It executes exactly what the CUDA sweep kernels execute -- tile gather by index bits, local ops, ext
conditions, per-group K' accumulators and the fused-group gradient finalisation -- so the planner and the
adjoint math can be checked against the oracle on a machine without a GPU.  It is NOT a fallback: the
product never imports it.
"""
import numpy as np

K_U1, K_D1, K_D1_EXT, K_CX, K_CX_EXT, K_CZ, K_CZ_EXT1, K_CZ_EXT2, K_SWAP = range(1, 10)
M_RX, M_RY, M_RZ, M_U = 1, 2, 3, 4
STEP_SWEEP, STEP_EXCHANGE = 0, 1


def member_matrix(mb, shared, batch_row, mats):
    if mb["kind"] == M_U:
        return np.asarray(mats[mb["slot"]], dtype=np.complex128)
    ang = batch_row[mb["slot"]] if mb["batch"] else shared[mb["slot"]]
    c, s = np.cos(ang / 2), np.sin(ang / 2)
    if mb["kind"] == M_RX:
        return np.array([[c, -1j * s], [-1j * s, c]])
    if mb["kind"] == M_RY:
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    return np.array([[np.exp(-1j * ang / 2), 0], [0, np.exp(1j * ang / 2)]])


PAULI = {M_RX: np.array([[0, 1], [1, 0]], dtype=np.complex128), M_RY: np.array([[0, -1j], [1j, 0]]),
         M_RZ: np.array([[1, 0], [0, -1]], dtype=np.complex128)}


def group_members(plan, g):
    return plan["members"][g["member_begin"]: g["member_begin"] + g["member_count"]]


def build_mats(plan, B, shared, batch, mats):
    """-> (mats_shared [Gs,2,2], mats_batch [B,Gb,2,2])"""
    ms = np.zeros((max(plan["n_groups_shared"], 1), 2, 2), np.complex128)
    mb = np.zeros((B, max(plan["n_groups_batch"], 1), 2, 2), np.complex128)
    for g in plan["groups"]:
        for b in range(B if g["batch"] else 1):
            U = np.eye(2, dtype=np.complex128)
            for m in group_members(plan, g):
                U = member_matrix(m, shared, batch[b] if batch is not None else None, mats) @ U
            if g["batch"]:
                mb[b, g["mat_index"]] = U
            else:
                ms[g["mat_index"]] = U
    return ms, mb


def _tile_index(sw, n_local):
    tb = sw["tile_bits"]
    m = len(tb)
    non = [b for b in range(n_local) if b not in tb]
    loc = np.arange(1 << m)
    off = np.zeros(1 << m, dtype=np.int64)
    for k, b in enumerate(tb):
        off |= ((loc >> k) & 1) << b
    tau = np.arange(1 << len(non))
    base = np.zeros(1 << len(non), dtype=np.int64)
    for k, b in enumerate(non):
        base |= ((tau >> k) & 1) << b
    return base, base[:, None] | off[None, :]


def _mat_for(op, ms, mb):
    idx = op["mat"] >> 1
    if op["mat"] & 1:
        return mb[:, idx]  # [B,2,2]
    return np.broadcast_to(ms[idx], (mb.shape[0], 2, 2))


def _apply_1q(t, a, M):
    """t: [B, nt, 2^m]; M: [B,2,2] applied on local bit a."""
    B, nt, sz = t.shape
    v = t.reshape(B, nt, sz >> (a + 1), 2, 1 << a)
    return np.einsum("bij,bthjl->bthil", M, v).reshape(B, nt, sz)


def _local_perm(m, fn):
    idx = np.arange(1 << m)
    return fn(idx)


def sweep_forward(plan, sw, state, ms, mb, rank=0):
    n_local = plan["n_local"]
    m = len(sw["tile_bits"])
    base, gidx = _tile_index(sw, n_local)
    gbase = base | (rank << n_local)
    t = state[:, gidx]  # [B, nt, 2^m]
    loc = np.arange(1 << m)
    for op in sw["ops"]:
        k, a, c = op["kind"], op["a"], op["c"]
        cond = ((gbase & op["ext_mask"]) == op["ext_mask"])[None, :, None]
        if k == K_U1:
            t = _apply_1q(t, a, _mat_for(op, ms, mb))
        elif k == K_D1:
            M = _mat_for(op, ms, mb)
            d = np.where(((loc >> a) & 1)[None, :] == 1, M[:, 1, 1][:, None], M[:, 0, 0][:, None])  # [B, 2^m]
            t = t * d[:, None, :]
        elif k == K_D1_EXT:
            M = _mat_for(op, ms, mb)
            bit = (gbase >> op["ext_bit"]) & 1  # [nt]
            d = np.where(bit[None, :] == 1, M[:, 1, 1][:, None], M[:, 0, 0][:, None])  # [B, nt]
            t = t * d[:, :, None]
        elif k == K_CX:
            src = np.where((loc >> c) & 1 == 1, loc ^ (1 << a), loc)
            t = t[:, :, src]
        elif k == K_CX_EXT:
            t = np.where(cond, t[:, :, loc ^ (1 << a)], t)
        elif k == K_CZ:
            sign = 1 - 2 * (((loc >> a) & 1) & ((loc >> c) & 1))
            t = t * sign[None, None, :]
        elif k == K_CZ_EXT1:
            sign = 1 - 2 * ((loc >> a) & 1)
            t = np.where(cond, t * sign[None, None, :], t)
        elif k == K_CZ_EXT2:
            t = np.where(cond, -t, t)
        elif k == K_SWAP:
            diff = ((loc >> a) & 1) != ((loc >> c) & 1)
            src = np.where(diff, loc ^ ((1 << a) | (1 << c)), loc)
            t = t[:, :, src]
        else:
            raise ValueError(k)
    out = state.copy()
    out[:, gidx] = t
    return out


def exchange(states, bits, n_local):
    """states: list over ranks of [B, 2^n_local]; swap rank bit j with local bit bits[j] (plan.h: Exchange)."""
    R = len(states)
    g = len(bits)
    assert R == 1 << g
    x = np.arange(1 << n_local)
    dest = np.zeros_like(x)  # B(x): the rank an amplitude goes to
    for j, b in enumerate(bits):
        dest |= ((x >> b) & 1) << j
    new = [s.copy() for s in states]
    for r in range(R):
        xr = x.copy()  # x with the exchanged bits replaced by r's bits
        for j, b in enumerate(bits):
            xr = (xr & ~(1 << b)) | (((r >> j) & 1) << b)
        for c in range(R):
            sel = dest == c
            new[r][:, x[sel]] = states[c][:, xr[sel]]
    return new


def emulate_forward(plan, state, shared, batch, mats, world=1):
    """state: [B, 2^n] complex (full, logical layout).  Returns the final full state in the PHYSICAL layout
    (apply final_pos to interpret) as a list over ranks when world > 1."""
    B = state.shape[0]
    ms, mb = build_mats(plan, B, shared, batch, mats)
    n_local = plan["n_local"]
    shards = [state[:, r << n_local:(r + 1) << n_local].astype(np.complex128) for r in range(world)]
    for st in plan["steps"]:
        if st["type"] == STEP_SWEEP:
            sw = plan["sweeps"][st["index"]]
            shards = [sweep_forward(plan, sw, s, ms, mb, rank=r) for r, s in enumerate(shards)]
        else:
            shards = exchange(shards, plan["exchanges"][st["index"]], n_local)
    return np.concatenate(shards, axis=1)


def probs_from_physical(plan, full):
    """P(q=0) using final_pos (what probs_finalize_kernel does)."""
    n = plan["n_qubits"]
    p = np.abs(full) ** 2
    idx = np.arange(p.shape[1])
    out = np.zeros((p.shape[0], n))
    for q in range(n):
        bit = plan["final_pos"][q]
        out[:, q] = p[:, ((idx >> bit) & 1) == 0].sum(axis=1)
    return out


def seed_probs(plan, full, g):
    n = plan["n_qubits"]
    idx = np.arange(full.shape[1])
    w = np.zeros((full.shape[0], full.shape[1]))
    for q in range(n):
        bit = plan["final_pos"][q]
        w += g[:, q][:, None] * (((idx >> bit) & 1) == 0)[None, :]
    return w * full


def sweep_backward(plan, sw, psi, lam, ms, mb, Ks, Kb, rank=0):
    """Mirror of sweep_backward_kernel: ops in reverse; K'[i][j] += sum psi_out[i] conj(lam_out[j])."""
    n_local = plan["n_local"]
    m = len(sw["tile_bits"])
    base, gidx = _tile_index(sw, n_local)
    gbase = base | (rank << n_local)
    tp, tl = psi[:, gidx], lam[:, gidx]
    loc = np.arange(1 << m)
    B = psi.shape[0]
    def add_K(kslot, sloc):  # sloc [B,3] = (sX, sY, sZ)
        ks = sw["kslots"][kslot]
        if ks["batch"]:
            Kb[:, ks["k_index"]] += sloc
        else:
            Ks[ks["k_index"]] += sloc.sum(axis=0)

    tdot_im = (np.conj(tl) * tp).imag.sum(axis=2)  # [B, nt], invariant under in-tile unitaries
    # flat plans carry the adjoint sweep's own linearisation (execution order); others walk `ops` in reverse
    bwd_ops = sw["ops_bwd"] if sw.get("ops_bwd") else list(reversed(sw["ops"]))
    for op in bwd_ops:
        k, a, c = op["kind"], op["a"], op["c"]
        cond = ((gbase & op["ext_mask"]) == op["ext_mask"])[None, :, None]
        if k == K_U1:
            M = _mat_for(op, ms, mb)
            if op["kslot"] >= 0:
                sz_ = 1 << m
                vp = tp.reshape(B, -1, sz_ >> (a + 1), 2, 1 << a)
                vl = tl.reshape(B, -1, sz_ >> (a + 1), 2, 1 << a)
                x, y, lx, ly = vp[:, :, :, 0], vp[:, :, :, 1], vl[:, :, :, 0], vl[:, :, :, 1]
                sX = (np.conj(lx) * y + np.conj(ly) * x).imag.sum(axis=(1, 2, 3))
                sY = ((np.conj(ly) * x).real - (np.conj(lx) * y).real).sum(axis=(1, 2, 3))
                sZ = (np.conj(lx) * x - np.conj(ly) * y).imag.sum(axis=(1, 2, 3))
                add_K(op["kslot"], np.stack([sX, sY, sZ], axis=1))
            Md = np.conj(np.swapaxes(M, 1, 2))
            tp, tl = _apply_1q(tp, a, Md), _apply_1q(tl, a, Md)
        elif k == K_D1:
            M = _mat_for(op, ms, mb)
            one = ((loc >> a) & 1) == 1
            if op["kslot"] >= 0:
                im = (np.conj(tl) * tp).imag
                sloc = np.zeros((B, 3))
                sloc[:, 2] = im[:, :, ~one].sum(axis=(1, 2)) - im[:, :, one].sum(axis=(1, 2))
                add_K(op["kslot"], sloc)
            d = np.where(one[None, :], np.conj(M[:, 1, 1])[:, None], np.conj(M[:, 0, 0])[:, None])
            tp, tl = tp * d[:, None, :], tl * d[:, None, :]
        elif k == K_D1_EXT:
            M = _mat_for(op, ms, mb)
            bit = ((gbase >> op["ext_bit"]) & 1) == 1  # [nt]
            if op["kslot"] >= 0:
                sloc = np.zeros((B, 3))
                sloc[:, 2] = tdot_im[:, ~bit].sum(axis=1) - tdot_im[:, bit].sum(axis=1)
                add_K(op["kslot"], sloc)
            d = np.where(bit[None, :], np.conj(M[:, 1, 1])[:, None], np.conj(M[:, 0, 0])[:, None])
            tp, tl = tp * d[:, :, None], tl * d[:, :, None]
        elif k == K_CX:
            src = np.where((loc >> c) & 1 == 1, loc ^ (1 << a), loc)
            tp, tl = tp[:, :, src], tl[:, :, src]
        elif k == K_CX_EXT:
            tp = np.where(cond, tp[:, :, loc ^ (1 << a)], tp)
            tl = np.where(cond, tl[:, :, loc ^ (1 << a)], tl)
        elif k == K_CZ:
            sign = 1 - 2 * (((loc >> a) & 1) & ((loc >> c) & 1))
            tp, tl = tp * sign, tl * sign
        elif k == K_CZ_EXT1:
            sign = 1 - 2 * ((loc >> a) & 1)
            tp = np.where(cond, tp * sign, tp)
            tl = np.where(cond, tl * sign, tl)
        elif k == K_CZ_EXT2:
            tp = np.where(cond, -tp, tp)
            tl = np.where(cond, -tl, tl)
        elif k == K_SWAP:
            diff = ((loc >> a) & 1) != ((loc >> c) & 1)
            src = np.where(diff, loc ^ ((1 << a) | (1 << c)), loc)
            tp, tl = tp[:, :, src], tl[:, :, src]
    po, lo = psi.copy(), lam.copy()
    po[:, gidx], lo[:, gidx] = tp, tl
    return po, lo


def finalize_grads(plan, B, shared, batch, mats, Ks, Kb, n_shared, n_batch_cols):
    """dtheta_k = a_k . s,  a_k = Pauli vector of V_k P_k V_k^+,  V_k = M_r ... M_{k+1}  (finalize_grads_kernel)."""
    gs = np.zeros(n_shared)
    gb = np.zeros((B, max(n_batch_cols, 1)))
    for g in plan["groups"]:
        if not g["has_param"]:
            continue
        for b in range(B if g["batch"] else 1):
            sv = Kb[b, g["k_index"]] if g["batch"] else Ks[g["k_index"]]
            V = np.eye(2, dtype=np.complex128)
            mem = group_members(plan, g)
            for mbr in reversed(mem):
                if mbr["kind"] != M_U:
                    A = V @ PAULI[mbr["kind"]] @ V.conj().T
                    val = A[1, 0].real * sv[0] + A[1, 0].imag * sv[1] + A[0, 0].real * sv[2]
                    if mbr["batch"]:
                        gb[b, mbr["slot"]] += val
                    else:
                        gs[mbr["slot"]] += val
                V = V @ member_matrix(mbr, shared, batch[b] if batch is not None else None, mats)
    return gs, gb


def emulate_backward(plan, psi_final, lam_final, shared, batch, mats, n_shared, n_batch_cols, world=1):
    """psi_final / lam_final: full physical-layout arrays [B, 2^n].  Returns (g_shared, g_batch, lam0, psi0)."""
    B = psi_final.shape[0]
    ms, mb = build_mats(plan, B, shared, batch, mats)
    Ks = np.zeros((max(plan["n_k_shared"], 1), 3))
    Kb = np.zeros((B, max(plan["n_k_batch"], 1), 3))
    n_local = plan["n_local"]
    ps = [psi_final[:, r << n_local:(r + 1) << n_local].astype(np.complex128) for r in range(world)]
    ls = [lam_final[:, r << n_local:(r + 1) << n_local].astype(np.complex128) for r in range(world)]
    for st in reversed(plan["steps"]):
        if st["type"] == STEP_SWEEP:
            sw = plan["sweeps"][st["index"]]
            for r in range(world):
                ps[r], ls[r] = sweep_backward(plan, sw, ps[r], ls[r], ms, mb, Ks, Kb, rank=r)
        else:
            ps, ls = exchange(ps, plan["exchanges"][st["index"]], n_local), exchange(ls, plan["exchanges"][st["index"]], n_local)
    gs, gb = finalize_grads(plan, B, shared, batch, mats, Ks, Kb, n_shared, n_batch_cols)
    return gs, gb, np.concatenate(ls, axis=1), np.concatenate(ps, axis=1)


# ---- thread-level model of the flat sweep kernels' shared-memory traffic (flat64.cuh / flat128.cuh) ----------------
def _ins0(x, p):
    return ((x >> p) << (p + 1)) | (x & ((1 << p) - 1))


def flat_thread_sets(sw, backward, packed, ext_on):
    """For every flat stage: (load, store) arrays [threads, amplitudes per thread] of the LOCAL tile indices a thread
    reads before / writes after the stage's register work, exactly as the kernels address them: thread g holds the
    amplitudes whose non-register bits spell g, loads through the inverse of the absorbed prefix CNOT maps and stores
    through the absorbed suffix maps (pk::absorb_maps).  ext_on: value of every out-of-tile control."""
    ops = sw["ops_bwd"] if backward else sw["ops"]
    stages = sw["stages_bwd"] if backward else sw["stages"]
    m = len(sw["tile_bits"])
    RB = 4 if packed else 3
    out = []
    g = np.arange(1 << (m - RB), dtype=np.int64)
    for st in stages:
        rb = st["regbits"]
        assert len(rb) == RB and rb == sorted(rb) and (not packed or rb[0] == 0)
        x = g << 1 if packed else g.copy()
        for p in (rb[1:] if packed else rb):
            x = _ins0(x, p)
        pat = np.zeros(1 << RB, dtype=np.int64)
        for j in range(1 << RB):
            for k, p in enumerate(rb):
                if (j >> k) & 1:
                    pat[j] |= 1 << p
        idx = x[:, None] | pat[None, :]

        def maps(v, lo, hi, reverse):
            v = v.copy()
            seq = range(hi - 1, lo - 1, -1) if reverse else range(lo, hi)
            for i in seq:
                o = ops[i]
                assert o["kind"] in (K_CX, K_CX_EXT)
                ctl = ((v >> o["c"]) & 1) if o["kind"] == K_CX else (1 if ext_on else 0)
                v = v ^ (ctl << o["a"])
            return v

        out.append((maps(idx, st["op_begin"], st["pre_end"], True), maps(idx, st["suf_begin"], st["op_end"], False)))
    return out


def check_flat_sync(sw, backward, packed):
    """The barriers the planner declared (Stage narrow_end / narrow_x) must cover every shared-memory hand-over between
    threads: returns the number of barriers by narrow code.  Raises AssertionError on a hazard."""
    stages = sw["stages_bwd"] if backward else sw["stages"]
    m = len(sw["tile_bits"])
    RB = 4 if packed else 3
    nthr = 1 << (m - RB)
    counts = [0, 0, 0, 0]
    for ext_on in (False, True):
        sets = flat_thread_sets(sw, backward, packed, ext_on)
        for s, st in enumerate(stages):
            ld, sto = sets[s]
            ne, nx = st["narrow_end"], st["narrow_x"]
            if ne or nx:
                assert nthr == 256, "narrow barriers are only defined for 256-thread tiles"
            # inside the stage: a thread may only overwrite slots it loaded itself, unless the stage has the extra
            # barrier after the loads -- then the hand-over must stay inside that barrier's thread group
            own = all(set(a) == set(b) for a, b in zip(ld.tolist(), sto.tolist()))
            if not st["xthread"]:
                assert own, f"stage {s}: cross-thread stores without the post-load barrier"
            else:
                grp = nthr >> nx
                for t0 in range(0, nthr, grp):
                    assert set(ld[t0:t0 + grp].ravel().tolist()) == set(sto[t0:t0 + grp].ravel().tolist()), \
                        f"stage {s}: cross-thread hand-over leaves its {grp}-thread barrier group"
            if s + 1 < len(stages):
                grp = nthr >> ne
                nld = sets[s + 1][0]
                for t0 in range(0, nthr, grp):
                    assert set(sto[t0:t0 + grp].ravel().tolist()) == set(nld[t0:t0 + grp].ravel().tolist()), \
                        f"stage {s} -> {s + 1}: data crosses the {grp}-thread barrier group"
            else:
                assert ne == 0, "the last stage must end with a CTA barrier (the tile store follows)"
            if not ext_on:
                counts[ne] += 1
                if st["xthread"]:
                    counts[nx] += 1
    return counts
