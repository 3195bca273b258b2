"""Tree output for whole batches: the decoded (peel, blens) of every sample / chain
come back from the device in ONE copy and are written as Newick on the host.

``tree_to_newick`` follows the reference's text format (dodonaphy/tree.py:207-267: tips as
``name:%.6f``, internal nodes ``(left,right):%.6f``, the root's two children at top level; unrooted
output folds the fake root into a trifurcation), ``save_trees`` appends the reference's sample lines
(dodonaphy/tree.py:168-198) for a batch.
"""
import os

import numpy as np
import torch


def _as_numpy(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def tree_to_newick(name_id, peel, blens, rooted=True):
    if not isinstance(name_id, dict):
        raise TypeError("name_id must be a dictionary")
    if len(name_id) == 0:
        return ";"
    if len(name_id) == 1:
        return f"({next(iter(name_id))});"
    label = {i: n for n, i in name_id.items()}
    peel = _as_numpy(peel)
    b = np.append(_as_numpy(blens).astype(np.float64), 1.0)
    n_tips = len(b) // 2 + 1
    text = {}

    def chunk(node):
        node = int(node)
        if node < n_tips:
            return f"{label[node]}:{b[node]:.6f}"
        return text[node]

    rows = n_tips - 1 if rooted else n_tips - 2
    for r in range(rows):
        left, right, parent = (int(v) for v in peel[r])
        if not rooted and r == rows - 1:
            third = int(min(peel[r + 1][:2]))
            # the reference also files the trifurcation under id n_tips before printing
            # (tree.py:249-255); reproduced so that the text is identical when that id is in use
            text[n_tips] = f"({chunk(left)},{chunk(right)},{chunk(third)}):{b[third + 1]:.6f}"
            return f"({chunk(left)},{chunk(right)},{chunk(third)};"
        if r == rows - 1:
            return f"({chunk(left)},{chunk(right)});"
        text[parent] = f"({chunk(left)},{chunk(right)}):{b[parent]:.6f}"
    return ";"


def save_trees(root_dir, filename, peels, blens, first_iteration, lnL, lnPr, name_id, lnQ=None, lnJac=None):
    """Append one sample line per tree of the batch to ``<root_dir>/<filename>.t``.
    peels (B,S-1,3), blens (B,2S-2), lnL/lnPr[/lnQ/lnJac] (B,): device or host arrays."""
    if root_dir is None:
        return
    peels, blens, lnL, lnPr = _as_numpy(peels), _as_numpy(blens), _as_numpy(lnL), _as_numpy(lnPr)
    lnQ = None if lnQ is None else _as_numpy(lnQ)
    lnJac = None if lnJac is None else _as_numpy(lnJac)
    lines = []
    for i in range(peels.shape[0]):
        tags = [f"&lnL={lnL[i]:f}", f"&lnPr={lnPr[i]:f}"]
        if lnQ is not None:
            tags.append(f"&lnQ={lnQ[i]:f}")
        if lnJac is not None:
            tags.append(f"&lnJac={lnJac[i]:f}")
        lines.append(f"\ttree STATE_{first_iteration + i} [{', '.join(tags)}] = [&U] {tree_to_newick(name_id, peels[i], blens[i])}\n")
    with open(os.path.join(root_dir, filename + ".t"), "a+", encoding="UTF-8") as fh:
        fh.writelines(lines)
