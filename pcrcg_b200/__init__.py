"""pcrcg_b200 -- B200 (sm_100a) implementation of PCR-CG's KPConv feature-extraction hot path.

  cpp_wrappers.cpp_subsampling.grid_subsampling / cpp_wrappers.cpp_neighbors.radius_neighbors
      drop-in modules for the reference's two CPython extensions (host buffers, NumPy results)
  dataloader   batch_grid_subsampling_kpconv, batch_neighbors_kpconv, collate_fn_descriptor, calibrate_neighbors
  blocks       KPConv, SimpleBlock, ResnetBottleneckBlock, UnaryBlock, max_pool, closest_pool, KPEncoder
  projection   Projection (3D -> 2D index lists) and the colour-feature scatter
  pipeline     FeaturePath: pyramid + encoder on stacked fragment pairs
  sharding     pair sharding across ranks and the NCCL gather of results
  ops          device-level operators (torch CUDA tensors) over the C ABI in include/pcrcg_b200.h

All compute happens in libpcrcg_b200.so (hand-written CUDA); there is no CPU fallback.
"""
__version__ = "0.1.0"
