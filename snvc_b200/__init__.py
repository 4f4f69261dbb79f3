"""snvc_b200 -- B200-native (sm_100a) implementation of SNVC's dense stereo-to-voxel hot path.

Package layout mirrors the part of the reference (`snvc/`) that is on the path:
  snvc_b200.extension.build_cost_volume   <-> snvc/extension/build_cost_volume/__init__.py
  snvc_b200.models.submodule              <-> snvc/models/submodule.py (3-D blocks)
  snvc_b200.models.vernier                <-> snvc/models/vernier.py   (instance hot path)
  snvc_b200.models.stereonet              <-> the (unshipped) global branch, SURVEY.md 3.4
  snvc_b200.csrc / libsnvc_b200.so        hand-written CUDA behind include/snvc_b200.h
"""
__version__ = "0.1.0"
