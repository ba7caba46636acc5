"""Graph -> SMILES / molfile stage with a PERSISTENT worker pool.

The reference's `convert_graph_to_smiles` (MolNexTR/chemical.py:960-975) forks a fresh 16-process
`multiprocessing.Pool` on every call and tears it down again, i.e. once per `predict_images` call
(MolNexTR/model.py:119): at the engine's rates that fork costs more than the model.  This module keeps the
pool alive across calls and keeps the reference's signature, chunking (`chunksize=128`), result order and
return value `(smiles_list, molblock_list, r_success)`.

The per-molecule function is NOT re-implemented here: it is RDKit chemistry
(`_convert_graph_to_smiles`, chemical.py:879-957) and RDKit is not available in this image, so there is
nothing to pin a restatement against.  `GraphPostProcessor()` imports the reference's function when
`MolNexTR.chemical` is importable and raises a clear error otherwise; any picklable function with the same
signature `(coords, symbols, edges[, image]) -> (smiles, molblock, success)` can be injected (tests do)."""
from __future__ import annotations

import atexit
import itertools
import multiprocessing
from typing import Callable, Optional, Sequence

import numpy as np


def _reference_worker() -> Callable:
    try:
        from MolNexTR.chemical import _convert_graph_to_smiles      # needs RDKit
        return _convert_graph_to_smiles
    except Exception as ex:      # pragma: no cover - depends on the installation
        raise RuntimeError("graph -> SMILES post-processing needs the reference's MolNexTR.chemical (RDKit); "
                           "it is outside the accelerated path") from ex


class GraphPostProcessor:
    """`post(coords, symbols, edges, images=None)` == `convert_graph_to_smiles(...)` of the reference, on a pool
    that is created once (lazily) and reused.  `num_workers <= 1` runs in-process, like the reference."""

    def __init__(self, worker: Optional[Callable] = None, num_workers: int = 16, start_method: Optional[str] = None):
        self.worker = worker if worker is not None else _reference_worker()
        self.num_workers = int(num_workers)
        self._ctx = multiprocessing.get_context(start_method) if start_method else multiprocessing
        self._pool = None
        atexit.register(self.close)

    def _ensure_pool(self):
        if self._pool is None:
            self._pool = self._ctx.Pool(self.num_workers)
        return self._pool

    def __call__(self, coords: Sequence, symbols: Sequence, edges: Sequence, images: Optional[Sequence] = None):
        if not (len(coords) == len(symbols) == len(edges)) or (images is not None and len(images) != len(coords)):
            raise ValueError("coords, symbols, edges (and images) must have one entry per molecule")
        args = zip(coords, symbols, edges) if images is None else zip(coords, symbols, edges, images)
        if len(coords) == 0:
            return (), (), float("nan")          # np.mean(()) in the reference (with a RuntimeWarning)
        if self.num_workers <= 1:
            results = list(itertools.starmap(self.worker, args))
        else:
            results = self._ensure_pool().starmap(self.worker, args, chunksize=128)
        smiles_list, molblock_list, success = zip(*results)
        return smiles_list, molblock_list, float(np.mean(success))

    def close(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool.join()
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
