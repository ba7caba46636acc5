"""`get_predictions` / `MolNexTRSingleton` with the reference's signatures (MolNexTR/molnextr.py:42-309).

The model path comes from $MOLNEXTR_B200_CHECKPOINT (the reference downloads its checkpoint through
pystow; there is no network here).  Errors propagate: the reference's "retry on CPU" path is removed
on purpose."""
from __future__ import annotations

import logging
import os

import torch

from .model import molnextr

logger = logging.getLogger("molnextr_singleton")
logger.setLevel(getattr(logging, os.environ.get("MOLNEXTR_DEBUG", "INFO").upper(), logging.INFO))


class MolNexTRSingleton:
    _instance = None
    _device = None
    _device_name = None

    @classmethod
    def get_device(cls):
        if cls._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("no CUDA device: molnextr_b200 has no CPU fallback")
            cls._device = torch.device("cuda")
            cls._device_name = f"NVIDIA GPU ({torch.cuda.get_device_name(0)})"
        return cls._device, cls._device_name

    @classmethod
    def get_instance(cls):
        if cls._instance is None:
            device, name = cls.get_device()
            path = os.environ.get("MOLNEXTR_B200_CHECKPOINT")
            if not path or not os.path.exists(path):
                raise FileNotFoundError("set MOLNEXTR_B200_CHECKPOINT to a MolNexTR checkpoint (.pth)")
            logger.info("Initializing MolNexTR (B200 engine) on %s", name)
            cls._instance = molnextr(path, device)
        return cls._instance


def get_predictions(imagepath: str, atoms_bonds: bool = False, smiles: bool = True, predicted_molfile: bool = False):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    model = MolNexTRSingleton.get_instance()
    predictions = model.predict_final_results(imagepath, return_atoms_bonds=atoms_bonds)
    result = {}
    if smiles:
        result["predicted_smiles"] = predictions["predicted_smiles"]
    if atoms_bonds:
        result["atom_sets"] = predictions["atom_sets"]
    if predicted_molfile:
        result["predicted_molfile"] = predictions["predicted_molfile"]
    result["device_info"] = MolNexTRSingleton.get_device()[1]
    end.record()
    torch.cuda.synchronize()
    result["prediction_time_seconds"] = start.elapsed_time(end) / 1000.0
    if not (smiles or atoms_bonds or predicted_molfile):
        return predictions
    return result
