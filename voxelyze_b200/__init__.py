"""voxelyze_b200 -- B200-native explicit dynamics step of Voxelyze (CVoxelyze::doTimeStep).

Layout (only what the hot path needs):
  csrc/       hand-written sm_100a CUDA kernels + the C-ABI implementation
  facade/     C++ mirror of the reference's public class API on top of the C-ABI
  lib/        built libvoxelyze_b200.so (git-ignored, travels with gpurun)
  capi.py     ctypes binding of include/voxelyze_b200.h
  scenarios.py  BASELINE.json configs as flat lattice descriptions
"""
from .capi import (Material, Sim, VxLib, VxError, load_product, MODEL_LINEAR, MODEL_BILINEAR,
                   MODEL_DATA, DOF_ALL)

__all__ = ["Material", "Sim", "VxLib", "VxError", "load_product", "MODEL_LINEAR",
           "MODEL_BILINEAR", "MODEL_DATA", "DOF_ALL"]
