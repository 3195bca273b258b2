"""Minimal Newick / NEXUS readers and a peel <-> Newick converter (host side).

The reference reads trees and alignments through DendroPy and converts them to its
``(peel, blens)`` representation in ``tree.dendropy_to_pb`` (dodonaphy/tree.py:69-122) and back in
``tree.tree_to_newick`` (dodonaphy/tree.py:207-267).  DendroPy is not a dependency here; these
helpers give the same representation:

* ``peel``  int64 (S-1, 3) rows ``[left, right, parent]`` in post-order, internal ids S.. in
  post-order, an unrooted (trifurcating) tree gets a fake root ``2S-2`` above the last internal
  node with a zero-length edge (tree.py:112-121);
* ``blens`` float64 (2S-2,), ``blens[i]`` = length of the branch above node ``i``.
"""
import re

import numpy as np


class _Nd:
    __slots__ = ("children", "label", "length")

    def __init__(self):
        self.children = []
        self.label = None
        self.length = 0.0


def parse_newick(text):
    """Parse one Newick string (comments in [] dropped) into a tree of _Nd."""
    text = re.sub(r"\[[^\]]*\]", "", text).strip()
    if text.endswith(";"):
        text = text[:-1]
    pos = 0

    def parse():
        nonlocal pos
        nd = _Nd()
        if text[pos] == "(":
            pos += 1
            while True:
                nd.children.append(parse())
                if text[pos] == ",":
                    pos += 1
                    continue
                if text[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad newick at {pos}")
        m = re.match(r"[^,():;]*", text[pos:])
        label = m.group(0).strip().strip("'\"")
        pos += m.end()
        nd.label = label or None
        if pos < len(text) and text[pos] == ":":
            m = re.match(r":\s*([0-9eE.+\-]+)", text[pos:])
            nd.length = float(m.group(1))
            pos += m.end()
        return nd

    return parse()


def newick_to_peel(text, taxon_labels, translate=None):
    """Newick -> (peel, blens) with the conventions of tree.dendropy_to_pb.
    ``taxon_labels``: list giving the tip index of every label; ``translate``: optional
    dict token -> label (NEXUS translate block)."""
    root = parse_newick(text)
    index = {lab: i for i, lab in enumerate(taxon_labels)}
    S = len(taxon_labels)
    peel = []
    blens = np.zeros(2 * S - 2, dtype=np.float64)
    counter = [S]

    def tip_id(nd):
        lab = nd.label
        if translate is not None and lab in translate:
            lab = translate[lab]
        if lab not in index:
            lab = lab.replace(" ", "_") if lab.replace(" ", "_") in index else lab.replace("_", " ")
        return index[lab]

    def visit(nd, is_root=False):
        if not nd.children:
            return tip_id(nd)
        ids = [visit(c) for c in nd.children]
        for c, i in zip(nd.children, ids):
            blens[i] = c.length
        if len(ids) == 2:
            me = counter[0]
            counter[0] += 1
            peel.append((ids[0], ids[1], me))
            return me
        if len(ids) == 3 and is_root:
            me = counter[0]
            counter[0] += 1
            peel.append((ids[0], ids[1], me))
            fake = 2 * S - 2
            peel.append((ids[2], me, fake))
            blens[me] = 0.0
            return fake
        raise ValueError("only binary trees (with an optional trifurcating root) are supported")

    visit(root, is_root=True)
    peel = np.array(peel, dtype=np.int64)
    if peel.shape[0] != S - 1:
        raise ValueError("tree does not span all taxa")
    return peel, blens


def peel_to_newick(peel, blens, tip_labels=None):
    """(peel, blens) -> Newick string rooted at the last parent (cf. tree.tree_to_newick)."""
    S = peel.shape[0] + 1
    if tip_labels is None:
        tip_labels = [f"T{i + 1}" for i in range(S)]
    chunks = {i: tip_labels[i] for i in range(S)}
    for left, right, parent in peel:
        left, right, parent = int(left), int(right), int(parent)
        chunks[parent] = (f"({chunks.pop(left)}:{float(blens[left]):.6e},"
                          f"{chunks.pop(right)}:{float(blens[right]):.6e})")
    return chunks[int(peel[-1, 2])] + ";"


def read_nexus_alignment(path):
    """Read a (possibly interleaved) NEXUS DNA matrix.  Returns (labels, sequences)."""
    labels, seqs = [], {}
    in_matrix = False
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        for raw in fh:
            line = re.sub(r"\[[^\]]*\]", "", raw).strip()
            if not line:
                continue
            low = line.lower()
            if not in_matrix:
                if low.startswith("matrix"):
                    in_matrix = True
                continue
            if line.startswith(";") or low.startswith("end"):
                break
            parts = line.rstrip(";").split()
            if len(parts) < 2:
                continue
            name, chunk = parts[0].strip("'"), "".join(parts[1:])
            if name not in seqs:
                seqs[name] = []
                labels.append(name)
            seqs[name].append(chunk)
            if line.endswith(";"):
                break
    return labels, ["".join(seqs[name]) for name in labels]


def read_nexus_tree(path):
    """Read the first tree of a NEXUS trees block.  Returns (newick, translate dict or None)."""
    translate = {}
    newick = None
    in_translate = False
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        for raw in fh:
            line = raw.strip()
            low = line.lower()
            if low.startswith("translate"):
                in_translate = True
                line = line[len("translate"):].strip()
                if not line:
                    continue
            if in_translate:
                done = line.endswith(";")
                for item in line.rstrip(";").split(","):
                    item = item.strip()
                    if item:
                        tok, lab = item.split(None, 1)
                        translate[tok] = lab.strip().strip("'")
                if done:
                    in_translate = False
                continue
            if low.startswith("tree "):
                newick = line.split("=", 1)[1].strip()
                break
    return newick, (translate or None)
