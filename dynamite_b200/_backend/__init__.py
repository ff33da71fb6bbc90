"""
Mirror of dynamite's compiled ``_backend`` package (``bbuild``, ``bsubspace``,
``bpetsc``) implemented over the C ABI of ``libdynamite_b200.so``.  Same module
names, function names and argument meaning as the Cython modules under
``/root/reference/src/dynamite/_backend/*.pyx`` so that dynamite's Python layer
can bind to this backend unchanged.
"""
