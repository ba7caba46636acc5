"""molnextr_b200 -- B200-native inference engine for the MolNexTR image -> atoms/bonds path.

Public surface mirrors the reference package (`MolNexTR/__init__.py`): `get_predictions`,
`MolNexTRSingleton`; `molnextr_b200.model.molnextr` mirrors `MolNexTR.model.molnextr`.
Importing the package does not load the CUDA library; creating an Engine does (and fails loudly
if it was not built)."""
__version__ = "0.1.0"

__all__ = ["get_predictions", "MolNexTRSingleton"]


def __getattr__(name):
    if name in __all__:
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
