"""quickchem_b200 — B200-native (sm_100a) OH hot path of GEOS-ESM/QuickChem.

The product is the C-ABI library ``libqcoh.so`` (``csrc/``, header ``include/qcoh.h``): the eleven XGBoost-named
symbols the reference's ``Shared/xgb_fortran_api.F90`` binds, plus the fused device-resident Run1 (``qcoh_*``).
This package only holds what the tests and ``bench.py`` need around it:

* ``capi``     ctypes binding of every symbol (no compute of its own; fails loudly without the library or a GPU)
* ``synth``    seeded synthetic met/chem fields and production-like boosters
* ``xgbmodel`` writers for XGBoost's legacy-binary / JSON / UBJSON model files
"""

__version__ = "0.1"
