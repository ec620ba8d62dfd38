#!/usr/bin/env python
"""Writes tests/golden/nrstubs_tree.npz from the reference's own fixture: the 36-taxon FBD tree with sampled ancestors that
examples/stubs/nrStubs.xml (= rjStubs.xml = stochasticStubs.xml) starts from, and its true parameters (nrStubs.xml:34-40).

The reference tree is parsed here as BEAST's TreeParser does with IsLabelledNewick="true", adjustTipHeights="false": node
heights = root height - distance from the root (tips are NOT moved to zero), taxon "k" is node k-1, internal nodes follow
(children before parents, root last).  README.md:144 gives the tree's summary statistics, checked below.
Run in the authoring container (it reads /root/reference):  python tests/golden/make_nrstubs_fixture.py"""
import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/examples/stubs/nrStubs.xml"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nrstubs_tree.npz")


def parse_newick(s):
    """-> list of nodes {label, length, children}, root last (post-order)"""
    s = s.strip().rstrip(";")
    pos = 0
    nodes = []

    def node():
        nonlocal pos
        children = []
        if s[pos] == "(":
            pos += 1
            while True:
                children.append(node())
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("bad newick at %d" % pos)
        m = re.match(r"[^:,()]*", s[pos:])
        label = m.group(0)
        pos += len(label)
        length = 0.0
        if pos < len(s) and s[pos] == ":":
            m = re.match(r":([0-9.eE+-]+)", s[pos:])
            length = float(m.group(1))
            pos += len(m.group(0))
        nodes.append({"label": label, "length": length, "children": children})
        return len(nodes) - 1

    root = node()
    assert root == len(nodes) - 1 and pos == len(s)
    return nodes


if __name__ == "__main__":
    x = ET.parse(REF).getroot()
    tree_el = [e for e in x.iter("stateNode") if e.get("id") == "tree"][0]
    newick = tree_el.get("newick")
    par = {e.get("id"): float(e.text) for e in x.iter("parameter") if e.get("id") in ("lambda", "r0", "samplingProportion")}
    nodes = parse_newick(newick)
    leaves = [i for i, nd in enumerate(nodes) if not nd["children"]]
    internal = [i for i, nd in enumerate(nodes) if nd["children"]]
    n = len(leaves)
    assert n == 36 and all(len(nodes[i]["children"]) == 2 for i in internal)
    nr = {}
    for i in leaves:
        nr[i] = int(nodes[i]["label"]) - 1          # taxon "k" -> node k-1
    for k, i in enumerate(internal):                 # post-order: children before parents, root last
        nr[i] = n + k
    N = len(nodes)
    parent = np.full(N, -1, np.int32)
    depth = np.zeros(N)
    def walk(i, d):
        depth[nr[i]] = d
        for c in nodes[i]["children"]:
            parent[nr[c]] = nr[i]
            walk(c, d + nodes[c]["length"])
    walk(len(nodes) - 1, 0.0)
    root_h = depth.max()
    height = root_h - depth
    # BEAST computes heights the same way (root height = deepest tip); sampled ancestors: zero-length leaves
    for c in range(N):                               # a zero-length leaf sits exactly at its parent's height
        if parent[c] >= 0 and nodes[[k for k, v in nr.items() if v == c][0]]["length"] == 0.0:
            height[c] = height[parent[c]]
    height[np.abs(height) < 1e-12] = 0.0
    extant = int(np.sum(height[:n] <= 1e-10))
    length = float(sum(nd["length"] for nd in nodes[:-1]))
    print("taxa %d extant %d non-extant %d height %.2f length %.2f" % (n, extant, n - extant, root_h, length))
    assert (extant, n - extant) == (20, 16) and abs(root_h - 6.53) < 0.006 and abs(length - 30.63) < 0.006   # README.md:144
    sa = int(np.sum((parent[:n] >= 0) & (height[:n] == height[np.maximum(parent[:n], 0)])))
    print("sampled ancestors:", sa)
    lam, r0, sp = par["lambda"], par["r0"], par["samplingProportion"]
    np.savez(OUT, height=height, parent=parent, lam=lam, r0=r0, sampling_proportion=sp,
             source="examples/stubs/nrStubs.xml:34-40 (tree) and :39-41 (parameters) of jordandouglas/GammaSpikeModel")
    print("wrote", OUT)
