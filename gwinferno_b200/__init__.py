"""gwinferno_b200 -- B200-native hierarchical population likelihood (fwd + VJP) behind the
call surface of FarrOutLab/GWInferno's model classes and ``hierarchical_likelihood``.

Layout:  csrc/        CUDA kernels + host plan builder + the C-ABI (libgwi.so, include/gwi.h)
         capi.py      ctypes binding of the C-ABI (fails loudly if the library is missing)
         spec.py      declarative model description shared with the C-ABI
         models.py    mirror of the reference's model classes (lazy weights)
         likelihood.py  ``hierarchical_likelihood`` / ``PopulationLikelihood`` front-end
         synthetic.py deterministic synthetic catalogs of the BASELINE.json shapes
"""

__version__ = "0.1.0"
