"""Multi-GPU host logic: LD blocks are independent (reference scr/dbslmmfit.cpp:193-213 touches only
index-b state), so N GPUs = N handles, each fitting its own shard; no collective on the data path.
Only the final beta vectors are gathered (torch.distributed, NCCL on GPUs / gloo in the CPU tests).
"""
import numpy as np

from . import _abi


def plan_owner(s_off, l_off, n_ref, n_ranks):
    """Block -> rank by the library's cost model + longest-processing-time-first (dbslmm_b200_plan_shards)."""
    lib = _abi.load()
    m_s = np.ascontiguousarray(np.diff(s_off), np.int32)
    m_l = None if l_off is None else np.ascontiguousarray(np.diff(l_off), np.int32)
    owner = np.zeros(m_s.size, np.int32)
    cost = np.zeros(n_ranks, np.float64)
    rc = lib.dbslmm_b200_plan_shards(m_s.size, m_s.ctypes.data, None if m_l is None else m_l.ctypes.data,
                                     int(n_ref), int(n_ranks), owner.ctypes.data, cost.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"dbslmm_b200_plan_shards failed ({rc})")
    return owner


def shard(w, owner, rank, pinned_alloc=None):
    """Compact per-rank problem: this rank's blocks (original order) and only their .bed rows.
    `w` holds bed[n_snp, pitch], CSR (s_off, s_pos, s_z, l_off, l_pos, l_z).  SNP rows used by the rank
    are renumbered 0..R-1 in first-use order so the rank uploads R * pitch bytes, not the whole panel."""
    bed = w["bed"]
    mine = np.where(owner == rank)[0]
    s_off = np.zeros(mine.size + 1, np.int32)
    l_off = np.zeros(mine.size + 1, np.int32)
    sp, lp, sz, lz = [], [], [], []
    for i, b in enumerate(mine):
        a = w["s_pos"][w["s_off"][b]:w["s_off"][b + 1]]
        c = w["l_pos"][w["l_off"][b]:w["l_off"][b + 1]] if w.get("l_off") is not None else np.zeros(0, np.int32)
        sp.append(a); lp.append(c)
        sz.append(w["s_z"][w["s_off"][b]:w["s_off"][b + 1]])
        lz.append(w["l_z"][w["l_off"][b]:w["l_off"][b + 1]] if w.get("l_off") is not None else np.zeros(0))
        s_off[i + 1] = s_off[i] + a.size
        l_off[i + 1] = l_off[i] + c.size
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    sp, lp = cat(sp, np.int64), cat(lp, np.int64)
    rows = np.unique(np.concatenate([sp, lp])) if (sp.size + lp.size) else np.zeros(0, np.int64)
    remap = np.full(bed.shape[0], -1, np.int64)
    remap[rows] = np.arange(rows.size)
    n_rows = max(int(rows.size), 1)
    if pinned_alloc is not None:
        sub = pinned_alloc((n_rows, bed.shape[1]))
    else:
        sub = np.zeros((n_rows, bed.shape[1]), np.uint8)
    if rows.size:
        sub[:rows.size] = bed[rows]
    return {"blocks": mine, "bed": sub, "s_off": s_off, "s_pos": remap[sp].astype(np.int32), "s_z": cat(sz, np.float64),
            "l_off": l_off, "l_pos": remap[lp].astype(np.int32), "l_z": cat(lz, np.float64)}


def gather_betas(w, owner, rank, world, beta_s, beta_l, dist):
    """Rank 0 reassembles block-major beta vectors from the per-rank results (all_gather_object keeps the
    test simple; sizes are a few MB genome-wide)."""
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(beta_s), np.asarray(beta_l)))
    full_s = np.zeros(int(w["s_off"][-1]))
    full_l = np.zeros(int(w["l_off"][-1])) if w.get("l_off") is not None else np.zeros(0)
    for r in range(world):
        bs, bl = parts[r]
        so = lo = 0
        for b in np.where(owner == r)[0]:
            ns = int(w["s_off"][b + 1] - w["s_off"][b])
            full_s[w["s_off"][b]:w["s_off"][b + 1]] = bs[so:so + ns]
            so += ns
            if full_l.size:
                nl = int(w["l_off"][b + 1] - w["l_off"][b])
                full_l[w["l_off"][b]:w["l_off"][b + 1]] = bl[lo:lo + nl]
                lo += nl
    return full_s, full_l
