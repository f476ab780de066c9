"""qcxms_b200 -- B200-native production-trajectory hot path of QCxMS (GFN2-xTB MD ensembles).

The product is the C-ABI shared library built from qcxms_b200/csrc (see include/qcxms_b200.h);
this package is the thin host-side mirror of the reference's Fortran interface for that path.
There is no CPU fallback: importing the API without the built CUDA library raises.
"""
from .api import (CidConfig, CidResult, Comm, Ensemble, MdConfig, MdResult, cid, cid_config, egrad_batch, fragment_structure, get_xtb_egrad, get_xtb_egrad_spec, gfn1_xtb, gfn2_xtb,
                  ipea1_xtb, lib, load_molecule, version)

from . import fragments, spectrum  # noqa: E402,F401

__all__ = ["fragments", "spectrum", "CidConfig", "CidResult", "Comm", "cid", "cid_config", "Ensemble", "MdConfig", "MdResult", "egrad_batch", "fragment_structure", "get_xtb_egrad", "get_xtb_egrad_spec", "gfn1_xtb",
           "gfn2_xtb", "ipea1_xtb", "lib", "load_molecule", "version"]
