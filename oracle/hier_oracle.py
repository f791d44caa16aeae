"""oracle/hier_oracle.py — numpy restatement of the reference's .hier file format and static cut (TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this; the product path never does).

Follows, under /root/reference/submodules/gaussianhierarchy:
  load()             HierarchyLoader::load        hierarchy_loader.cpp:26-128
  write()            HierarchyWriter::write       hierarchy_writer.cpp:27-118
  expand_to_target() Traversal::expandToTarget    traversal.cpp:14-38  (recExpand)
Node = 7 x int32 (depth, parent, start, count_leafs, count_merged, start_children, count_children; types.h:47-56);
HalfNode = 3 x int32 (parent, start, start_children) + 4 x int16 (depth, count_children, count_leafs, count_merged;
types.h:58-64).  Half conversions are IEEE round-to-nearest-even (half.hpp 2.2.0, HALF_ROUND_STYLE 1) = numpy's.

PINNED: tests/golden/hier_ref_*.hier were written by the reference's own HierarchyWriter and
tests/golden/hier_ref_expected.npz holds what its HierarchyLoader / expandToTarget return for them
(tests/golden/make_hier_golden.py, through oracle/_ref/ref_hier_io.so).
"""
import numpy as np

HALF_NODE = np.dtype([("parent", "<i4"), ("start", "<i4"), ("start_children", "<i4"), ("dccc", "<i2", (4,))])


def load(path):
    """-> dict(pos [P,3], shs [P,48], alphas [P], scales [P,3], rot [P,4], nodes [N,7] i32, boxes [N,2,4], compressed)."""
    raw = np.fromfile(path, dtype=np.uint8)
    P = int(raw[:4].view("<i4")[0])
    off = 4

    def take(count, dtype):
        nonlocal off
        nbytes = count * np.dtype(dtype).itemsize
        out = raw[off:off + nbytes].view(dtype)
        off += nbytes
        return out

    if P >= 0:  # hierarchy_loader.cpp:43-67
        pos = take(P * 3, "<f4").reshape(P, 3).copy()
        rot = take(P * 4, "<f4").reshape(P, 4).copy()
        scales = take(P * 3, "<f4").reshape(P, 3).copy()
        alphas = take(P, "<f4").copy()
        shs = take(P * 48, "<f4").reshape(P, 48).copy()
        N = int(take(1, "<i4")[0])
        nodes = take(N * 7, "<i4").reshape(N, 7).copy()
        boxes = take(N * 8, "<f4").reshape(N, 2, 4).copy()
        return dict(pos=pos, shs=shs, alphas=alphas, scales=scales, rot=rot, nodes=nodes, boxes=boxes, compressed=False)
    P = -P  # hierarchy_loader.cpp:68-127
    pos = take(P * 3, "<f4").reshape(P, 3).copy()
    rot = take(P * 4, "<f2").astype(np.float32).reshape(P, 4)
    scales = take(P * 3, "<f2").astype(np.float32).reshape(P, 3)
    alphas = take(P, "<f2").astype(np.float32)
    shs = take(P * 48, "<f2").astype(np.float32).reshape(P, 48)
    N = int(take(1, "<i4")[0])
    hn = take(N, HALF_NODE)
    nodes = np.empty((N, 7), np.int32)
    nodes[:, 0] = hn["dccc"][:, 0]
    nodes[:, 1] = hn["parent"]
    nodes[:, 2] = hn["start"]
    nodes[:, 3] = hn["dccc"][:, 2]
    nodes[:, 4] = hn["dccc"][:, 3]
    nodes[:, 5] = hn["start_children"]
    nodes[:, 6] = hn["dccc"][:, 1]
    boxes = take(N * 8, "<f2").astype(np.float32).reshape(N, 2, 4)
    return dict(pos=pos, shs=shs, alphas=alphas, scales=scales, rot=rot, nodes=nodes, boxes=boxes, compressed=True)


def file_bytes(pos, shs, opacities, log_scales, rotations, nodes, boxes, compressed=True):
    """The exact bytes HierarchyWriter::write produces (hierarchy_writer.cpp:44-117)."""
    f32 = lambda a: np.ascontiguousarray(a, dtype="<f4").reshape(-1)  # noqa: E731
    P, N = len(pos), len(nodes)
    nodes = np.ascontiguousarray(nodes, dtype="<i4").reshape(N, 7)
    if not compressed:
        parts = [np.array([P], "<i4"), f32(pos), f32(rotations), f32(log_scales), f32(opacities), f32(shs),
                 np.array([N], "<i4"), nodes.reshape(-1), f32(boxes)]
    else:
        if N and (nodes[:, [0, 6, 3, 4]] > 32000).any():
            raise RuntimeError("Would lose information!")
        def f16(a):
            with np.errstate(over="ignore"):  # values beyond 65504 become inf, as half.hpp rounds them
                return f32(a).astype("<f2")
        hn = np.zeros(N, HALF_NODE)
        hn["parent"], hn["start"], hn["start_children"] = nodes[:, 1], nodes[:, 2], nodes[:, 5]
        hn["dccc"][:, 0], hn["dccc"][:, 1] = nodes[:, 0], nodes[:, 6]
        hn["dccc"][:, 2], hn["dccc"][:, 3] = nodes[:, 3], nodes[:, 4]
        parts = [np.array([-P], "<i4"), f32(pos), f16(rotations), f16(log_scales), f16(opacities), f16(shs),
                 np.array([N], "<i4"), hn, f16(boxes)]
    return b"".join(np.ascontiguousarray(p).tobytes() for p in parts)


def write(path, *args, **kw):
    with open(path, "wb") as f:
        f.write(file_bytes(*args, **kw))


def expand_to_target(nodes, target):
    """recExpand from node 0 (traversal.cpp:14-38)."""
    out = []

    def rec(i):
        depth, _parent, start, leafs, merged, first_child, children = (int(v) for v in nodes[i])
        out.extend(range(start, start + leafs))
        if depth <= target:
            out.extend(range(start + leafs, start + leafs + merged))
        else:
            for c in range(children):
                rec(first_child + c)
    rec(0)
    return np.asarray(out, np.int32)


def synthetic_hierarchy(n_leaves=40, seed=0, branching=(2, 4)):
    """A small valid hierarchy in the reference's node layout: a random tree whose leaves own one Gaussian each
    (count_leafs = 1) and whose inner nodes own one merged Gaussian (count_merged = 1); depth counts up from the leaves'
    level as the reference's builders emit it (the root has the largest depth)."""
    rng = np.random.default_rng(seed)
    # build top-down: list of (children ids); then assign depth = height above the deepest leaf
    children = {0: []}
    frontier, n_nodes, leaves = [0], 1, 0
    while frontier and leaves + len(frontier) < n_leaves:
        node = frontier.pop(0)
        k = int(rng.integers(branching[0], branching[1] + 1))
        ids = list(range(n_nodes, n_nodes + k))
        n_nodes += k
        children[node] = ids
        for c in ids:
            children[c] = []
        frontier.extend(ids)
    height = {}

    def h(i):
        if i not in height:
            height[i] = 0 if not children[i] else 1 + max(h(c) for c in children[i])
        return height[i]
    h(0)
    # children of a node must be contiguous: they are (allocated consecutively above)
    nodes = np.zeros((n_nodes, 7), np.int32)
    nxt = 0
    parent = {0: -1}
    for i in range(n_nodes):
        for c in children[i]:
            parent[c] = i
    for i in range(n_nodes):
        leaf = not children[i]
        nodes[i] = (height[i], parent[i], nxt, 1 if leaf else 0, 0 if leaf else 1,
                    children[i][0] if children[i] else -1, len(children[i]))
        nxt += 1
    P = nxt
    g = dict(pos=rng.normal(0, 5, (P, 3)).astype(np.float32), shs=rng.normal(0, 0.3, (P, 48)).astype(np.float32),
             alphas=rng.uniform(0.01, 0.99, P).astype(np.float32), scales=rng.normal(-3, 0.7, (P, 3)).astype(np.float32),
             rot=rng.normal(0, 1, (P, 4)).astype(np.float32))
    lo = g["pos"][nodes[:, 2]] - 0.3
    hi = g["pos"][nodes[:, 2]] + 0.3
    boxes = np.zeros((n_nodes, 2, 4), np.float32)
    boxes[:, 0, :3], boxes[:, 1, :3] = lo, hi
    boxes[:, 0, 3] = rng.uniform(0.1, 1.0, n_nodes)
    g.update(nodes=nodes, boxes=boxes)
    return g
