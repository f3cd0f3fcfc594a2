"""B200-native H|psi> (OpPsi) on the Smolyak type-4 sparse grid -- drop-in for ElVibRot-TnumTana's
sub_TabOpPsi_FOR_SGtype4 path.  The package name contains a hyphen (it is fixed by the project
layout); import it with ``importlib.import_module("elvibrot-tnumtana_b200")`` or through the
``evr_sg4_b200`` alias module at the repo root.
"""
from . import algebra, distributed, lib, primitives, sg4, workloads  # noqa: F401
from .lib import EvrSg4Error, build  # noqa: F401
from .sg4 import (Basis_L_TO_n, EvrStop, Init_TypeOp, OpGrid, ParamOp, ParamOp10, ParamPsi, SG4Basis, SG4Transforms,  # noqa: F401
                  level_sizes, sub_OpPsi, sub_scaledOpPsi, sub_TabOpPsi, sub_TabOpPsi_FOR_SGtype4)

__all__ = ["algebra", "distributed", "lib", "primitives", "sg4", "workloads", "build", "EvrSg4Error", "EvrStop", "Basis_L_TO_n",
           "Init_TypeOp", "OpGrid", "ParamOp", "ParamOp10", "ParamPsi", "SG4Basis", "SG4Transforms", "level_sizes", "sub_OpPsi",
           "sub_scaledOpPsi", "sub_TabOpPsi", "sub_TabOpPsi_FOR_SGtype4"]
