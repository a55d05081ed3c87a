"""newtonnet.data of the reference: only what the energy/force path and its callers use is mirrored - RadiusGraph
(the transform utils/ase_interface.py:12 imports; here the neighbour search runs inside the model, csrc/nbr.cu) and the
PyG-free data side of newtonnet_b200.data (extended-xyz reader, collation, MolecularStatistics)."""
from newtonnet_b200.layers.representations import RadiusGraph      # noqa: F401
from newtonnet_b200.data import *                                  # noqa: F401,F403


def __getattr__(name):
    if name in ('MolecularDataset', 'MolecularInMemoryDataset', 'parse_train_test'):
        raise ImportError(f'newtonnet.data.{name} (torch_geometric dataset machinery) is outside the B200 energy/force path; '
                          f'use newtonnet_b200.data.read_extxyz / collate / molecular_statistics')
    raise AttributeError(name)
