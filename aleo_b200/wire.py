"""Compressed G1 wire format on the device (include/aleo_b200.h ``aleo_b200_g1_{de,}compress_dev``): snarkVM's
CanonicalSerialize / CanonicalDeserialize of ``G1Affine`` -- the 48-byte points inside proofs and key files (SURVEY.md
section 8f rank 3; format pinned by the reference's proof string wasm/src/programs/transaction.rs:100)."""
from __future__ import annotations

from . import _lib


def g1_decompress_dev(compressed_t, affine_stride: int = 104):
    """(n * 48) uint8 CUDA tensor -> ((n * stride) uint8 CUDA tensor of affine points, number of invalid encodings)"""
    import torch

    lib = _lib.get_lib()
    n = compressed_t.numel() * compressed_t.element_size() // 48
    out = torch.empty(max(n, 1) * affine_stride, dtype=torch.uint8, device=compressed_t.device)
    with torch.cuda.device(compressed_t.device):
        bad = lib.check(lib.g1_decompress_dev(out.data_ptr(), affine_stride, compressed_t.data_ptr(), n,
                                              torch.cuda.current_stream().cuda_stream), "aleo_b200_g1_decompress_dev")
    return out[: n * affine_stride], bad


def g1_compress_dev(affine_t, n: int, affine_stride: int = 104):
    """n affine points (stride 104 / 96) -> (n * 48) uint8 CUDA tensor"""
    import torch

    lib = _lib.get_lib()
    out = torch.empty(max(n, 1) * 48, dtype=torch.uint8, device=affine_t.device)
    with torch.cuda.device(affine_t.device):
        lib.check(lib.g1_compress_dev(out.data_ptr(), affine_t.data_ptr(), affine_stride, n,
                                      torch.cuda.current_stream().cuda_stream), "aleo_b200_g1_compress_dev")
    return out[: n * 48]
