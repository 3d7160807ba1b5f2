"""``con``: contract a labelled network given as tensors + label pairs
(reference: tncon.py:4-159).  Pure host bookkeeping; every contraction is a
``contract`` / ``trace`` call and therefore a libtnb kernel."""
from . import tensor as tsr

__all__ = ["con"]


def _split_args(args):
    """Tensors and pairs may come loose or in lists (tncon.py:68-81)."""
    tensors, pairs = [], []
    for x in args:
        if isinstance(x, list):
            (tensors if isinstance(x[0], tsr.Tensor) else pairs).extend(x)
        elif isinstance(x, tsr.Tensor):
            tensors.append(x)
        else:
            pairs.append(x)
    return tensors, pairs


def con(*args):
    """Contract the network.  Edges inside one tensor are traced first; the
    remaining edges are contracted tensor pair by tensor pair in the order the
    pairs were given (all edges between the same two tensors in one call);
    disconnected components are multiplied together at the end."""
    tensors, pairs = _split_args(args)
    tensors = [t.copy() for t in tensors]

    ends = [lab for pair in pairs for lab in pair]
    if len(set(ends)) != len(ends):
        raise ValueError("Index found in more than one contraction pair.")
    home = {}
    for i, t in enumerate(tensors):
        for lab in t.labels:
            if lab in home:
                raise ValueError("Index label " + lab + " found in two tensors."
                                 " Tensors must have unique index labelling.")
            home[lab] = i

    loops, edges, keys = [], [], []
    for c in pairs:
        i, j = home[c[0]], home[c[1]]
        if i == j:
            loops.append(c)
            continue
        key = (min(i, j), max(i, j))
        if key in keys:
            # the reference looks the tuple up in the order the pair was written
            # (tncon.py:112-113), so a repeated edge written "backwards" is a ValueError
            k = keys.index((i, j))
            if not isinstance(edges[k][0], list):
                edges[k] = [[edges[k][0]], [edges[k][1]]]
            edges[k][0].append(c[0])
            edges[k][1].append(c[1])
        else:
            edges.append(list(c))
            keys.append(key)

    for c in loops:
        tensors[home[c[0]]].trace(c[0], c[1])

    root = list(range(len(tensors)))
    for c in edges:
        first = (c[0][0], c[1][0]) if isinstance(c[0], list) else (c[0], c[1])
        d, e = home[first[0]], home[first[1]]
        if d == e:  # an earlier contraction already merged the two tensors
            tensors[d].trace(c[0], c[1])
            continue
        if d < e:
            tensors[d] = tsr.contract(tensors[d], tensors[e], c[0], c[1])
            root[e] = d
        else:
            tensors[e] = tsr.contract(tensors[e], tensors[d], c[1], c[0])
            root[d] = e
        keep = min(d, e)
        for lab in tensors[keep].labels:
            home[lab] = keep
    return tsr.tensor_product(*[tensors[root.index(x)] for x in set(root)])
