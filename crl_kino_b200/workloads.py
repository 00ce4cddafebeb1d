"""Synthetic workloads of BASELINE.json configs 4 and 5:
dense random-obstacle maps and query batches.  Deterministic (numpy Generator
seeds), no reference data needed."""
import numpy as np

WALLS = np.array([[-22, -22, 22, -20], [-22, 20, 22, 22], [-22, -20, -20, 20], [20, -20, 22, 20]],
                 dtype=np.float32)   # crl_kino/utils/__init__.py:18-21


def _aabb_intersect(a, b):
    """crl_kino/utils/__init__.py aabb_intersect_aabb rule, boxes as [l, b, r, t]."""
    return not (a[2] < b[0] or a[0] > b[2] or a[3] < b[1] or a[1] > b[3])


def random_rectangles(rng, n_obs=8, size_range=(3.0, 20.0), min_gap=3.0, bounds=(-20.0, 20.0)):
    """The generation rule of the reference's obs_generate (utils/__init__.py:26-56)."""
    out = []
    lo, hi = bounds
    guard = 0
    while len(out) < n_obs and guard < 100000:
        guard += 1
        l, b = rng.uniform(lo, hi - 5, 2)
        w, h = rng.uniform(size_range[0], size_range[1], 2)
        r, t = l + w, b + h
        if not 40 <= w * h <= 150:
            continue
        if r > hi or t > hi:
            continue
        box = (l, b, r, t)
        if all(not _aabb_intersect(box, (o[0] - min_gap, o[1] - min_gap, o[2] + min_gap, o[3] + min_gap))
               for o in out):
            out.append(box)
    return np.array(out, dtype=np.float64)


def dense_map(n_cells=10000, cell=0.25, seed=0):
    """cfg4 map: obs_generate-style rectangles tiled into `cell`-sized RectObs cells
    until n_cells of them exist, plus the 4 boundary walls -> float32 [n_cells + 4, 4]."""
    rng = np.random.default_rng(seed)
    cells = []
    guard = 0
    while len(cells) < n_cells and guard < 64:
        guard += 1
        for (l, b, r, t) in random_rectangles(rng, n_obs=8):
            nx = max(1, int(np.floor((r - l) / cell)))
            ny = max(1, int(np.floor((t - b) / cell)))
            xs = l + cell * np.arange(nx)
            ys = b + cell * np.arange(ny)
            gx, gy = np.meshgrid(xs, ys, indexing="ij")
            c = np.stack([gx.ravel(), gy.ravel(), gx.ravel() + cell, gy.ravel() + cell], 1)
            cells.append(c)
            if sum(len(x) for x in cells) >= n_cells:
                break
        if sum(len(x) for x in cells) >= n_cells:
            break
    cells = np.concatenate(cells)[:n_cells]
    return np.ascontiguousarray(np.concatenate([cells.astype(np.float32), WALLS]))


def random_states(rng, n, bounds=(-20.0, 20.0), moving=False):
    s = np.zeros((n, 5))
    s[:, 0] = rng.uniform(bounds[0], bounds[1], n)
    s[:, 1] = rng.uniform(bounds[0], bounds[1], n)
    s[:, 2] = rng.uniform(-np.pi, np.pi, n)
    if moving:
        s[:, 3] = rng.uniform(-0.1, 1.0, n)
        s[:, 4] = rng.uniform(-np.pi, np.pi, n)
    return s


def free_states(clearance_fn, rng, n, min_clearance=0.5, moving=False, batch=4096):
    """Rejection-sample n states whose clearance exceeds min_clearance (the
    reference's own reset rule, differential_gym.py:229,238).  clearance_fn maps
    [B, 5] float64 -> [B] clearance (numpy)."""
    out = []
    have = 0
    while have < n:
        c = random_states(rng, batch, moving=moving)
        ok = clearance_fn(c) > min_clearance
        out.append(c[ok])
        have += int(ok.sum())
    return np.concatenate(out)[:n]


def query_pairs(clearance_fn, n, seed=1, min_dist=5.0):
    """cfg4 queries: (start, goal) pairs at rest, clearance > 0.5, ||dxy|| > min_dist."""
    rng = np.random.default_rng(seed)
    starts, goals = [], []
    have = 0
    while have < n:
        s = free_states(clearance_fn, rng, n, 0.5)
        g = free_states(clearance_fn, rng, n, 0.5)
        s[:, 2:] = 0.0
        g[:, 2:] = 0.0
        ok = np.linalg.norm(s[:, :2] - g[:, :2], axis=1) > min_dist
        starts.append(s[ok]); goals.append(g[ok])
        have += int(ok.sum())
    return np.concatenate(starts)[:n], np.concatenate(goals)[:n]


def sample_stream(state_bounds, goal, n, seed, goal_sample_rate=5):
    """Pre-drawn equivalent of RRT.sample (rrt.py:146-154): n candidate states,
    each the goal with probability (goal_sample_rate + 1) / 101, else uniform."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(state_bounds[:, 0], state_bounds[:, 1], (n, 5))
    pick_goal = rng.integers(0, 101, n) <= goal_sample_rate
    u[pick_goal] = np.asarray(goal, dtype=np.float64)
    return u


DENSE_DWA = dict(dt=0.2, predict_time=1.0, v_res=0.4 / 63, w_res=0.4 * np.pi / 63)   # 64 x 64 nominal grid


def device_sample_stream(state_bounds, goals, n, seed, goal_sample_rate=5):
    """Host mirror of the planner kernel's own sample generator (crlk_sample_stream, include/crlk.h):
    [Q, n, 5] float64, bit-identical to what crlk_plan_dwa draws when it is given no streams.
    Counter based: u01(seed, q, i, c) = (splitmix64_finalise(seed ^ q*K1 ^ i*K2 ^ (c+1)*K3) >> 11) * 2^-53;
    component c of draw i is lo + u * (hi - lo); draw i is the goal iff floor(101 * u01(.., 5)) <= rate
    (RRT.sample, rrt.py:146-154: random.randint(0, 100) > goal_sample_rate picks the uniform branch)."""
    goals = np.asarray(goals, dtype=np.float64).reshape(-1, 5)
    Q = len(goals)
    sb = np.asarray(state_bounds, dtype=np.float64)
    u64 = np.uint64
    q = np.arange(Q, dtype=u64)[:, None, None]
    i = np.arange(n, dtype=u64)[None, :, None]
    c = np.arange(1, 7, dtype=u64)[None, None, :]
    with np.errstate(over="ignore"):
        z = u64(int(seed) & 0xFFFFFFFFFFFFFFFF) ^ (q * u64(0x100000001B3)) ^ (i * u64(0x9E3779B97F4A7C15)) \
            ^ (c * u64(0xD6E8FEB86659FD93))
        z ^= z >> u64(30)
        z *= u64(0xBF58476D1CE4E5B9)
        z ^= z >> u64(27)
        z *= u64(0x94D049BB133111EB)
        z ^= z >> u64(31)
    u = (z >> u64(11)).astype(np.float64) * 1.1102230246251565e-16
    out = sb[:, 0] + u[:, :, :5] * (sb[:, 1] - sb[:, 0])
    pick_goal = np.floor(u[:, :, 5] * 101.0).astype(np.int64) <= goal_sample_rate
    out[pick_goal] = np.repeat(goals[:, None, :], n, 1)[pick_goal]
    return out


# ---------------------------------------------------------------- cfg4 batches (bench / parity)

def clearance_raster(rects, pad, res=0.125, lo=-22.0, hi=22.0):
    """Boolean raster [n, n]: True where some obstacle rectangle comes within `pad` of the raster
    cell (conservative: every rectangle is grown by pad and rounded outwards to whole cells)."""
    gx = np.arange(lo, hi + res, res)
    occ = np.zeros((len(gx), len(gx)), dtype=bool)
    for l, b, r, t in np.asarray(rects, dtype=np.float64):
        i0, i1 = np.searchsorted(gx, [l - pad, r + pad])
        j0, j1 = np.searchsorted(gx, [b - pad, t + pad])
        occ[max(i0 - 1, 0):i1 + 1, max(j0 - 1, 0):j1 + 1] = True
    return gx, occ


def raster_free(gx, occ, xy):
    ii = np.clip(np.searchsorted(gx, xy[:, 0]) - 1, 0, len(gx) - 1)
    jj = np.clip(np.searchsorted(gx, xy[:, 1]) - 1, 0, len(gx) - 1)
    return ~occ[ii, jj]


def cfg4_workload(Q=1024, tree_nodes=1024, n_batches=25, seed=100, cells=10000, cand=96):
    """BASELINE config 4, generated on the host with numpy / scipy only so that every arm of
    bench.py (GPU, CPU port) and the tests work on the SAME data: the dense map, Q trees of
    `tree_nodes` free nodes at rest, and n_batches x Q accepted samples -- valid states (rrt.py:66)
    whose nearest tree node lies in [1, 20) (rrt.py:69).  Free / valid are decided by conservative
    rasters: no obstacle within 1.5 m of a tree node (clearance > 0.5, the reference's reset rule,
    differential_gym.py:229) and none within 0.95 m of a sample (> footprint circumradius + 0.05:
    valid at any heading).  Only ~5 % of the valid candidates have no tree node within 1 m, so each
    (query, batch) slot draws `cand` candidates and keeps the first accepted one."""
    from scipy.spatial import cKDTree
    rects = dense_map(n_cells=cells)
    rng = np.random.default_rng(seed)
    gx, occ_node = clearance_raster(rects, 1.5)
    _, occ_sample = clearance_raster(rects, 0.95)
    nodes = []
    have = 0
    while have < Q * tree_nodes:
        c = random_states(rng, 1 << 18)
        c = c[raster_free(gx, occ_node, c)]
        nodes.append(c)
        have += len(c)
    trees = np.concatenate(nodes)[:Q * tree_nodes].reshape(Q, tree_nodes, 5)
    trees[:, :, 3:] = 0.0
    batches = np.zeros((n_batches, Q, 5))
    for q in range(Q):
        kd = cKDTree(trees[q, :, :2])
        # one generator per (query, batch): a batch does not depend on how many batches are drawn
        gens = [np.random.default_rng([seed, q, b]) for b in range(n_batches)]
        need = np.ones(n_batches, dtype=bool)
        while need.any():
            todo = np.flatnonzero(need)
            c = np.stack([random_states(gens[b], cand, moving=True) for b in todo]).reshape(-1, 5)
            ok = raster_free(gx, occ_sample, c)
            d = np.full(len(c), np.inf)
            d[ok] = kd.query(c[ok, :2])[0]
            ok &= (d >= 1.0 + 1e-9) & (d < 20.0 - 1e-9)
            ok = ok.reshape(len(todo), cand)
            first = ok.argmax(axis=1)
            hit = ok.any(axis=1)
            batches[todo[hit], q] = c.reshape(len(todo), cand, 5)[hit, first[hit]]
            need[todo[hit]] = False
    return dict(rects=rects, trees=trees, batches=batches)
