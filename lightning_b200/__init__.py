"""lightning_b200 — B200-native batched secp256k1 ECDSA / BIP-340 verification engine behind
Core Lightning's bitcoin/signature.h surface.  See DESIGN.md."""
from .engine import (KIND_ECDSA33, KIND_ECDSA_XY, KIND_SCHNORR, KEY_SIZE, EngineError, SigVerifier, SvTx, load_library)

__all__ = ["KIND_ECDSA33", "KIND_ECDSA_XY", "KIND_SCHNORR", "KEY_SIZE", "EngineError", "SigVerifier", "SvTx", "load_library"]
