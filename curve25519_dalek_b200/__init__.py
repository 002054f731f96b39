"""curve25519_dalek_b200 -- host-side mirror of the reference's multiscalar / batch-verify API over
the C ABI of libdalek_b200.so (include/dalek_b200.h).

The product path is the CUDA library; there is no CPU fallback.  Importing this package never
touches the CPU checker used by the tests.  Names follow the reference:
  EdwardsPoint.vartime_multiscalar_mul / optional_multiscalar_mul / multiscalar_mul
      (curve25519-dalek/src/traits.rs:78-262, src/edwards.rs:966-1031)
  RistrettoPoint.multiscalar_mul / vartime_multiscalar_mul (src/ristretto.rs:964-994)
  VartimeEdwardsPrecomputation / VartimeRistrettoPrecomputation (traits.rs:290-406, edwards.rs:1038-1076)
  verify_batch (ed25519-dalek/src/batch.rs:146-251) and its SignatureError values.
"""
from .engine import (Engine, MultiEngine, EngineError, EdwardsPoint, RistrettoPoint, SignatureError, verify_batch, default_engine,
                     library_path, load_library, POINTS_COMPRESSED, POINTS_EXTENDED, POINTS_RISTRETTO,
                     VartimeEdwardsPrecomputation, VartimeRistrettoPrecomputation)

__all__ = ["Engine", "MultiEngine", "EngineError", "EdwardsPoint", "RistrettoPoint", "SignatureError", "verify_batch", "default_engine",
           "library_path", "load_library", "POINTS_COMPRESSED", "POINTS_EXTENDED", "POINTS_RISTRETTO",
           "VartimeEdwardsPrecomputation", "VartimeRistrettoPrecomputation"]
