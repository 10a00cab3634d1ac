"""rna-sequence-diff-patch_b200 — B200-native (sm_100a) weighted Wagner–Fischer engine.

A drop-in for the hot path of plsakr/rna-sequence-diff-patch: StringEditDistance.py
(distance, edit script, patch) and IRMethods.wf_score / search_collection.  Python talks to
hand-written CUDA through the C ABI in include/rsd.h (librsd.so, loaded with ctypes).
There is no CPU fallback: compute calls raise RsdError without a CUDA device.

Layout:
  csrc/        CUDA kernels + the C ABI
  _lib.py      ctypes binding of every symbol in include/rsd.h
  encoding.py  symbols <-> 4-bit codes, packed batches
  engine.py    Engine: costs, batched distance / script / patch / search
  sed.py       the StringEditDistance.py module surface on top of Engine
  ir.py        IRMethods.wf_score / search_collection / top-k on top of Engine
  ingest.py    FASTA / SeqXML -> packed database (T->U, X->N like the reference's importers)
  eswire.py    edit-script JSON export/import, packed <-> dict scripts, reverse on packed scripts
  dist_search.py  database search sharded over the GPUs of one box (one all_gather)
  dist_pairs.py   pair batches sharded over the GPUs of one box (balanced by cells, no exchange)
  dropin/      modules importable as `StringEditDistance` / `IRMethods` + cost files
"""
from ._lib import RsdError, load_library, library_path  # noqa: F401
from .encoding import SYMBOLS, PackedSeqs, encode, decode, pack  # noqa: F401
from .engine import Engine, MultiEngine, get_engine, get_search_engine  # noqa: F401

__all__ = ["Engine", "MultiEngine", "get_engine", "get_search_engine", "RsdError", "PackedSeqs", "SYMBOLS", "encode", "decode", "pack",
           "load_library", "library_path"]
