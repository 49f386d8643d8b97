"""Neighbour-table build and file format (``entrypoints/generate_codebook.py`` in the reference).

The reference computes ``torch.cdist`` + ``topk(N-1)`` on the VQ codebook and stores ``top_{N-1}_indices.npy`` as
uint16 ``[N, N-1]`` (:53-65); the EA models ``np.load`` it (ea_model_llamagen.py:143, ea_model_anole.py:142,
ea_model_lumina_mgpt.py:321).  Here the table is built on the GPU by ``lantern_build_neighbors`` (exact order:
squared L2 distance accumulated in fp64, ties broken by id) and kept as int32 ``[N, K]`` with K = lantern_k + 1
columns — all the verification step ever reads.  ``save_reference_format`` / ``load_neighbor_table`` keep wire
compatibility with existing ``ckpts/*/vq_distances`` files.
"""
from __future__ import annotations

import argparse
import os
from typing import Optional

import numpy as np
import torch

from . import _abi


_workspaces = {}


def build_neighbor_table(embedding: torch.Tensor, k: Optional[int] = None, return_route: bool = False):
    """embedding: [N, d] codebook (``vq_model.quantize.embedding.weight``) -> int32 [N, K] on the same device.
    ``return_route``: also return the device int32[2] record (1 = tensor-core route for every row / 2 = some rows by the
    all-fp64 kernel, and how many)."""
    lib = _abi.load()
    if embedding.dim() != 2:
        raise ValueError("embedding must be [N, d]")
    if not embedding.is_cuda:
        raise ValueError("build_neighbor_table runs on the GPU: move the codebook to a CUDA device")
    E = embedding.detach().to(torch.float32).contiguous()
    N, d = E.shape
    K = N - 1 if k is None else int(k)
    out = torch.empty(N, K, dtype=torch.int32, device=E.device)
    need = int(lib.lantern_build_neighbors_workspace_bytes(N, d, K))
    key = str(E.device)
    work = _workspaces.get(key)                      # caller-owned scratch, kept between calls (up to 1 GiB at N = 16384)
    if work is None or work.numel() < need:
        work = _workspaces[key] = torch.empty(max(need, 1), dtype=torch.uint8, device=E.device)
    route = torch.zeros(2, dtype=torch.int32, device=E.device)
    _abi.check(lib.lantern_build_neighbors(E.data_ptr(), N, d, K, out.data_ptr(), work.data_ptr(), work.numel(),
                                           route.data_ptr(), torch.cuda.current_stream(E.device).cuda_stream))
    return (out, route) if return_route else out


def save_reference_format(table: torch.Tensor, save_path: str) -> str:
    """Write ``top_{K}_indices.npy`` as uint16, the layout the reference loaders expect (generate_codebook.py:60-65)."""
    t = table.detach().cpu().numpy()
    if t.max() >= 65536:
        raise ValueError("codebook ids do not fit uint16")
    os.makedirs(save_path, exist_ok=True)
    path = os.path.join(save_path, f"top_{t.shape[1]}_indices.npy")
    np.save(path, t.astype(np.uint16))
    return path


def load_neighbor_table(path: str, cols: Optional[int] = None, device=None) -> torch.Tensor:
    """Load a reference-format (uint16) or native (int32) table; keep the first ``cols`` columns; int32 on ``device``."""
    arr = np.load(path, mmap_mode="r")
    if cols is not None:
        arr = arr[:, :cols]
    t = torch.from_numpy(np.ascontiguousarray(arr).astype(np.int32))
    return t.to(device) if device is not None else t


def parse_args():
    parser = argparse.ArgumentParser(description="Generate codebook")
    parser.add_argument("--model", type=str, default="lumina_mgpt", help="Model type; names the default save path")
    parser.add_argument("--save_path", type=str, default=None, help="Path to save the codebook")
    parser.add_argument("--embedding", type=str, required=True,
                        help=".npy / .pt file with the VQ codebook weights [N, d] (loading the VQ checkpoints is outside "
                             "this package; export `vq_model.quantize.embedding.weight` once)")
    parser.add_argument("--k", type=int, default=None, help="columns to keep (default N-1, the reference's file format)")
    return parser


def run_generate_codebook(args) -> str:
    """entrypoints/generate_codebook.py:15-65 with the codebook given as a file."""
    save_path = args.save_path or f"ckpts/{args.model}/vq_distances"
    if args.embedding.endswith(".npy"):
        emb = torch.from_numpy(np.load(args.embedding))
    else:
        emb = torch.load(args.embedding, map_location="cpu")
    table = build_neighbor_table(emb.cuda(), args.k)
    return save_reference_format(table, save_path)


if __name__ == "__main__":
    print(run_generate_codebook(parse_args().parse_args()))
