"""Drop-in modules for the reference's ``structureLibs`` package, hot path only.

    from waterorderlib_b200.structureLibs import water_properties as wp      # was: import water_properties as wp
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    from waterorderlib_b200.structureLibs import waterlib as wl               # was: the f2py module

Same function names, positional order, keyword names and defaults, return types (numpy in -> numpy out;
torch CUDA tensors in -> torch CUDA tensors out).  Everything is computed by libwol.so's sm_100a kernels;
there is no CPU fallback.
"""
