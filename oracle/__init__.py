"""ORACLE -- test infrastructure, not product code.

CPU restatements of the reference (Nicholasli1995/SNVC) algorithms on the dense
stereo-to-voxel hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; ``snvc_b200`` (the product) never does.

Modules
-------
cost_volume   numpy restatement + ctypes binding of ``cost_volume.c``
              (reference: snvc/extension/build_cost_volume/src/BuildCostVolume_cuda.cu:15-98,152-205)
grid_sample   numpy restatement of the bilinear / trilinear zero-padded sampler the
              reference calls (vernier.py:339-340 -> torch ATen GridSampler.h:27-36)
blocks        plain-torch restatement of convbn_3d / hourglass / hourglass_downsample_16
              (snvc/models/submodule.py:32-50,85-168,170-268)
global_branch restated global trunk + frustum lift (blocks pinned
              by submodule.py, wiring restated from the DSGN lineage, README.md:68)
instance_branch ROI voxel sampling + refinement 3-D CNN (vernier.py:323-360,414-438)
grid_proj     numpy float64 restatement of refinementDataset._generate_grid_proj (N1)
iou3d_nms     scalar fp32 restatement of the rotated BEV IoU / NMS (N4) + exact float64 clipper
"""
