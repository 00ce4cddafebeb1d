"""The reference's on-disk formats, readable and writable without the reference installed
(SURVEY.md section 8f, row 4):

* obstacle maps -- pickled ``list[crl_kino.utils.rigid.RectObs]`` (data/obstacles/*.pkl and
  obs_list_list.pkl, a list of such lists)            <-> flat float32 ``[M, 4]`` arrays;
* benchmark positions -- pickled list of 5-vectors (examples/benchmark_position_test_env*.pkl);
* planning results -- ``{'runtime': array, 'path_len': array}`` (examples/benchmark.py:38-43,
  :86-88; runtime -1 marks a query that did not reach its goal);
* trees and paths -- pickled ``planner.node_list`` (objects of ``RRT.Node``: state, path,
  parent) and pickled ``path`` (list of 5-vectors), examples/test.py:219-220.

Reading maps the reference's class paths onto this package's containers; writing emits the
reference's class paths, so the reference's own ``pickle.load`` accepts the files.
"""
import contextlib
import io
import pickle
import sys
import types

import numpy as np

from .rigid import CircleRobot, RectObs, RectRobot, rects_of

_REF_RIGID = "crl_kino.utils.rigid"
_REF_RRT = "crl_kino.planner.rrt"


class _Reader(pickle.Unpickler):
    def find_class(self, module, name):
        if module == _REF_RIGID:
            return {"RectObs": RectObs, "RectRobot": RectRobot, "CircleRobot": CircleRobot}[name]
        if module.startswith("crl_kino.planner") and name.endswith("Node"):
            from ..planner.rrt import RRT
            return RRT.Node
        return super().find_class(module, name)


def _load(src):
    if isinstance(src, (bytes, bytearray)):
        return _Reader(io.BytesIO(bytes(src))).load()
    with open(src, "rb") as f:
        return _Reader(f).load()


@contextlib.contextmanager
def _reference_paths():
    """While active, ``crl_kino.utils.rigid.RectObs`` and ``crl_kino.planner.rrt.RRT`` resolve to
    stand-in classes carrying the reference's module / qualified names, which is all pickle
    records.  Modules that already exist (the real reference is importable) are left alone."""
    made = []

    def module(name):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            made.append(name)
        return sys.modules[name]

    for pkg in ("crl_kino", "crl_kino.utils", "crl_kino.planner"):
        module(pkg)
    rigid, rrt = module(_REF_RIGID), module(_REF_RRT)
    if not hasattr(rigid, "RectObs"):
        rigid.RectObs = type("RectObs", (), {"__module__": _REF_RIGID})
    if not hasattr(rrt, "RRT"):
        node = type("Node", (), {"__module__": _REF_RRT, "__qualname__": "RRT.Node"})
        rrt.RRT = type("RRT", (), {"__module__": _REF_RRT, "Node": node})
    try:
        yield rigid.RectObs, rrt.RRT.Node
    finally:
        for name in made:
            sys.modules.pop(name, None)


def _dump(obj, dst):
    data = pickle.dumps(obj, protocol=4)
    if dst is None:
        return data
    with open(dst, "wb") as f:
        f.write(data)
    return data


# ---- obstacle maps -----------------------------------------------------------------

def load_obstacles(src):
    """Reference obstacle pickle (path or bytes) -> list[RectObs] (or list of lists)."""
    return _load(src)


def load_obstacles_flat(src):
    """-> float32 [M, 4] (left, bottom, right, top), or a list of such arrays for obs_list_list.pkl."""
    obs = _load(src)
    if len(obs) and isinstance(obs[0], (list, tuple)):
        return [rects_of(o) for o in obs]
    return rects_of(obs)


def save_obstacles(rects, dst=None, color="black"):
    """float32 [M, 4] (or list[RectObs]) -> a pickle the reference loads as list[RectObs]."""
    flat = rects_of(rects)
    with _reference_paths() as (ref_obs, _):
        out = []
        for r in flat:
            o = ref_obs.__new__(ref_obs)
            o.__dict__.update(color=color, rect=np.array(r.reshape(2, 2), dtype=np.float32))   # rigid.py:126, :146
            out.append(o)
        return _dump(out, dst)


# ---- benchmark positions / results ---------------------------------------------------

def load_positions(src):
    """examples/benchmark_position_test_env*.pkl -> float64 [K, 5]."""
    return np.asarray(_load(src), dtype=np.float64).reshape(-1, 5)


def benchmark_pairs(positions):
    """The ordered (start, goal) pairs of examples/benchmark.py:56-59."""
    return [(i, j) for i in range(len(positions)) for j in range(len(positions)) if j != i]


def planning_results(reach_exactly, planning_time, path_len_steps, dt=0.2):
    """The dict examples/benchmark.py builds (:38-43, :69-75, :86-87): every query appends its
    runtime (or -1 when the goal was not reached); path_len only gets reached queries."""
    data = {"runtime": [], "path_len": []}
    for ok, t, n in zip(reach_exactly, planning_time, path_len_steps):
        if ok:
            data["path_len"].append(dt * (n - 1))
            data["runtime"].append(t)
        else:
            data["runtime"].append(-1)
    return {k: np.array(v) for k, v in data.items()}


def save_results(data, fname):
    """benchmark.py:88 (fname without the extension)."""
    with open(fname + ".pkl", "wb") as f:
        pickle.dump({k: np.asarray(v) for k, v in data.items()}, f)


def load_results(src):
    return _load(src)


# ---- trees and paths ----------------------------------------------------------------

def save_tree(node_list, dst=None):
    """planner.node_list -> pickle of reference ``RRT.Node`` objects (test.py:219)."""
    with _reference_paths() as (_, ref_node):
        twin = {}
        out = []
        for n in node_list:                      # parents precede children in node_list
            m = ref_node.__new__(ref_node)
            m.__dict__.update(state=np.array(n.state, dtype=np.float64),
                              path=[np.array(p, dtype=np.float64) for p in n.path],
                              parent=twin.get(id(n.parent)))
            twin[id(n)] = m
            out.append(m)
        return _dump(out, dst)


def load_tree(src):
    """-> list of RRT.Node (state, path, parent links restored)."""
    return _load(src)


def tree_arrays(node_list):
    """node_list -> (states [n, 5], parent index [n] with -1 for the root)."""
    index = {id(n): i for i, n in enumerate(node_list)}
    states = np.array([n.state for n in node_list], dtype=np.float64).reshape(-1, 5)
    parents = np.array([-1 if n.parent is None else index[id(n.parent)] for n in node_list], dtype=np.int32)
    return states, parents


def save_path(path, dst=None):
    """planning()'s return value (list of 5-vectors) as test.py:220 pickles it."""
    return _dump([np.array(p, dtype=np.float64) for p in path], dst)


def load_path(src):
    return np.asarray(_load(src), dtype=np.float64).reshape(-1, 5)
