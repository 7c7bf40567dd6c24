"""B200-native iterative decoding hot path of SophieLangdon27/LDPC_SPARC.

Modules mirror the reference's layout for this path only:
  ldpc      -> ldpc_jossy/py/ldpc.py      (``code`` object, BP decoder)
  sparc     -> sparc_public/sparc.py      (sparc_encode / sparc_decode, Ab / Az, AMP)
  sparc_new -> sparc_sophie/sparc_new.py  (dense-design AMP, AMP+BP decoders)
All decoding runs in libldpcsparc_b200.so (hand-written sm_100a CUDA behind the
C-ABI of include/ldpc_sparc_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
