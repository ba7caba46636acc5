"""CPU: the persistent-pool plumbing of the graph -> SMILES stage (row f-3) against the reference's
`convert_graph_to_smiles` contract (MolNexTR/chemical.py:960-975): argument zipping with / without images, result
order, `(smiles_list, molblock_list, r_success)`, in-process mode for num_workers <= 1 -- with a stand-in for the RDKit
function (RDKit is not in this image)."""
import os

import pytest

from molnextr_b200.postprocess import GraphPostProcessor


def fake_convert(coords, symbols, edges, image=None):
    """Same signature and return shape as chemical._convert_graph_to_smiles."""
    n_bonds = sum(1 for i in range(len(symbols)) for j in range(i + 1, len(symbols)) if edges[i][j])
    smiles = ".".join(symbols) + f"|{n_bonds}|{os.getpid()}"
    molblock = f"{len(coords)} atoms" + ("" if image is None else f" {image[0]}x{image[1]}")
    return smiles, molblock, len(symbols) % 2 == 0


def _mols(n):
    coords = [[[0.1 * k, 0.2 * k] for k in range(i % 4 + 1)] for i in range(n)]
    symbols = [[f"C{k}" for k in range(i % 4 + 1)] for i in range(n)]
    edges = [[[int(a != b) for b in range(i % 4 + 1)] for a in range(i % 4 + 1)] for i in range(n)]
    return coords, symbols, edges


@pytest.mark.parametrize("workers", [1, 3])
def test_order_and_success_rate_match_the_reference_contract(workers):
    coords, symbols, edges = _mols(300)                       # more than one chunk of 128
    with GraphPostProcessor(worker=fake_convert, num_workers=workers) as post:
        smiles, molblocks, rate = post(coords, symbols, edges)
        assert len(smiles) == len(molblocks) == 300
        for i in range(300):
            want, wantb, _ = fake_convert(coords[i], symbols[i], edges[i])
            assert smiles[i].rsplit("|", 1)[0] == want.rsplit("|", 1)[0] and molblocks[i] == wantb     # order preserved
        assert abs(rate - sum(len(s) % 2 == 0 for s in symbols) / 300) < 1e-12
        images = [(10 + i, 20 + i) for i in range(300)]
        _, molblocks2, _ = post(coords, symbols, edges, images=images)
        assert molblocks2[7].endswith("17x27")


def test_pool_is_created_once_and_reused():
    coords, symbols, edges = _mols(400)
    with GraphPostProcessor(worker=fake_convert, num_workers=2) as post:
        pids = set()
        for _ in range(3):                                    # the reference forks a new pool on every call
            smiles, _, _ = post(coords, symbols, edges)
            pids |= {s.rsplit("|", 1)[1] for s in smiles}
        assert 1 <= len(pids) <= 2 and str(os.getpid()) not in pids


def test_argument_errors_and_missing_rdkit_are_loud():
    with GraphPostProcessor(worker=fake_convert, num_workers=1) as post:
        with pytest.raises(ValueError):
            post([[]], [[], []], [[]])
    try:
        import rdkit  # noqa: F401
        have = True
    except Exception:
        have = False
    if not have:
        with pytest.raises(RuntimeError, match="RDKit"):
            GraphPostProcessor()
