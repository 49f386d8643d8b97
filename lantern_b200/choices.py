"""Static draft-tree shapes for LANTERN++ / EAGLE-1 drafting.

Same six trees the reference ships in ``models/drafters/choices.py:1-31`` (tree data,
originally from the EAGLE / Medusa projects), stored as digit strings: ``"012"`` is the
path "child 0 of root -> its child 1 -> its child 2".  ``tree(name)`` returns the
list-of-paths form the reference API uses (``tree_choices=``).
"""
from __future__ import annotations

from typing import Dict, List

_PATHS: Dict[str, str] = {
    "mc_sim_7b_63":
        "0 1 2 3 00 01 02 10 11 20 21 30 000 001 002 010 011 020 021 100 0000 0001 0002 00000 00001",
    "mc_sim_7b_63_balanced":
        "0 1 2 00 01 02 10 11 12 20 21 000 001 002 010 011 100 101 110 111 0000 0001 0002 00000 00001",
    "naive_extend_57":
        "0 1 2 3 4 00 01 02 03 10 11 12 20 21 22 30 31 40 000 001 002 003 010 011 012 020 021 022 030 031 "
        "100 101 110 200 0000 0001 0002 0003 0010 0011 0012 0020 0021 0030 0100 0101 0110 0200 00000 00001 "
        "00002 00010 00011 00100 00101 00110 00200",
    "medusa_2_7b_63":
        "0 1 2 3 4 5 6 7 8 9 00 01 02 03 04 05 06 07 08 09 10 11 12 13 14 20 21 30 31 40 50 60 70 000 001 "
        "002 003 004 005 006 007 008 010 011 012 013 020 021 030 040 050 100 101 102 110 200 0000 0001 0002 "
        "0003 0010 0020 0100",
    "reverse_balanced_25":
        "0 1 2 00 01 10 20 000 001 002 010 011 100 0000 0001 0002 0010 0011 00000 00001 00002 00003 00010 "
        "00011 00012",
    "chain": "0 00 000 0000 00000",
}

NAMES = tuple(_PATHS)
# BASELINE configs[2] asks for "tree sizes 10-80 nodes": beyond the six shipped shapes, `synth_<nodes>_<seed>` names a
# deterministic random prefix-closed tree (children numbered consecutively, at most 10 per node, depth <= 5)
SYNTH_NAMES = ("synth_10_1", "synth_24_2", "synth_40_3", "synth_80_4")


def synth_tree(n_nodes: int, seed: int, max_depth: int = 5, max_children: int = 10) -> List[List[int]]:
    state = (seed * 2654435761 + 12345) & 0xFFFFFFFF

    def rnd(n: int) -> int:
        nonlocal state
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        return (state >> 8) % n

    paths: List[tuple] = [()]
    kids = {(): 0}
    while len(paths) - 1 < n_nodes:
        open_ = [p for p in paths if len(p) < max_depth and kids[p] < max_children]
        # favour first children (deep, narrow trees like the shipped ones): two draws, keep the parent with fewer kids
        a, b = open_[rnd(len(open_))], open_[rnd(len(open_))]
        par = a if kids[a] <= kids[b] else b
        child = par + (kids[par],)
        kids[par] += 1
        kids[child] = 0
        paths.append(child)
    return [list(p) for p in paths[1:]]


def tree(name: str) -> List[List[int]]:
    if name.startswith("synth_"):
        _, n, seed = name.split("_")
        return synth_tree(int(n), int(seed))
    return [[int(c) for c in path] for path in _PATHS[name].split()]


def __getattr__(name: str):
    if name in _PATHS:
        return tree(name)
    raise AttributeError(name)
