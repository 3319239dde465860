"""Drop-in replacements for the reference's two CPython extension modules, same import paths
relative to this package (``datasets/dataloader.py:5-6``):

    import pcrcg_b200.cpp_wrappers.cpp_subsampling.grid_subsampling as cpp_subsampling
    import pcrcg_b200.cpp_wrappers.cpp_neighbors.radius_neighbors as cpp_neighbors
"""
