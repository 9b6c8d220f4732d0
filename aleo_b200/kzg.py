"""Resident SRS and the KZG10 commitment path.

snarkVM's ``KZG10::commit(powers, polynomial, hiding_bound, rng)`` (snarkvm-algorithms 0.14.5
src/polycommit/kzg10/mod.rs) is ``VariableBase::msm(&powers.powers_of_beta_g[..d + 1], coeffs.to_bigint())``
with the SAME bases for every commitment.  ``ResidentSRS`` keeps those bases on the GPU, expanded once
(``aleo_b200_srs_create``); ``KZG10.commit`` mirrors the non-hiding commitment and returns the 48-byte
compressed G1 the proof serialises.  (The hiding term is a second, small MSM over
``powers_of_beta_times_gamma_g`` with a random polynomial -- a second ResidentSRS and a point addition on the
caller's side; the randomness belongs to snarkVM's prover and is out of this path.)
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from .msm import _host_ptr, PROJECTIVE_BYTES, AFFINE_STRIDE_RUST


class ResidentSRS:
    def __init__(self, handle, lib):
        self._h = handle
        self._lib = lib

    @classmethod
    def from_host(cls, bases, affine_stride: int = AFFINE_STRIDE_RUST):
        lib = _lib.get_lib()
        ptr, nbytes, _keep = _host_ptr(bases)
        h = C.c_void_p()
        lib.check(lib.srs_create(C.byref(h), ptr, nbytes // affine_stride, affine_stride), "aleo_b200_srs_create")
        return cls(h, lib)

    @classmethod
    def from_device(cls, bases_t, n: int, affine_stride: int = AFFINE_STRIDE_RUST):
        import torch

        lib = _lib.get_lib()
        h = C.c_void_p()
        with torch.cuda.device(bases_t.device):
            lib.check(lib.srs_create_dev(C.byref(h), bases_t.data_ptr(), n, affine_stride, torch.cuda.current_stream().cuda_stream),
                      "aleo_b200_srs_create_dev")
        return cls(h, lib)

    def info(self) -> dict:
        n, c, w, b = C.c_size_t(), C.c_int(), C.c_int(), C.c_size_t()
        self._lib.check(self._lib.srs_info(self._h, C.byref(n), C.byref(c), C.byref(w), C.byref(b)), "aleo_b200_srs_info")
        return {"n": n.value, "window_bits": c.value, "windows": w.value, "device_bytes": b.value}

    def msm(self, scalars) -> bytes:
        """sum_i s_i P_i over the first len(scalars) bases; canonical 32-byte scalars on the host"""
        sp, sbytes, _keep = _host_ptr(scalars)
        out = C.create_string_buffer(PROJECTIVE_BYTES)
        self._lib.check(self._lib.srs_msm(self._h, C.cast(out, C.c_void_p), sp, sbytes // 32), "aleo_b200_srs_msm")
        return out.raw

    def msm_dev(self, scalars_t, n_used: int, out=None):
        import torch

        if out is None:
            out = torch.empty(PROJECTIVE_BYTES, dtype=torch.uint8, device=scalars_t.device)
        with torch.cuda.device(scalars_t.device):
            self._lib.check(self._lib.srs_msm_dev(self._h, out.data_ptr(), scalars_t.data_ptr(), n_used,
                                                  torch.cuda.current_stream().cuda_stream), "aleo_b200_srs_msm_dev")
        return out

    def launches(self, n_used: int) -> int:
        return self._lib.srs_msm_launches(self._h, n_used)

    def close(self):
        if self._h is not None:
            self._lib.srs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KZG10:
    @staticmethod
    def commit(srs: ResidentSRS, coeffs_montgomery) -> bytes:
        """non-hiding KZG10::commit: Montgomery-form Fr coefficients (host) -> 48-byte compressed G1"""
        cp, cbytes, _keep = _host_ptr(coeffs_montgomery)
        out = C.create_string_buffer(48)
        srs._lib.check(srs._lib.kzg_commit(srs._h, C.cast(out, C.c_void_p), cp, cbytes // 32), "aleo_b200_kzg_commit")
        return out.raw

    @staticmethod
    def commit_dev(srs: ResidentSRS, coeffs_t, n_coeffs: int, out=None):
        import torch

        if out is None:
            out = torch.empty(48, dtype=torch.uint8, device=coeffs_t.device)
        with torch.cuda.device(coeffs_t.device):
            srs._lib.check(srs._lib.kzg_commit_dev(srs._h, out.data_ptr(), coeffs_t.data_ptr(), n_coeffs,
                                                   torch.cuda.current_stream().cuda_stream), "aleo_b200_kzg_commit_dev")
        return out

    @staticmethod
    def commit_batch_dev(srs: ResidentSRS, coeff_tensors, out=None):
        """`count` (<= 64) non-hiding commitments against the same resident powers in ONE launch sequence
        (aleo_b200_kzg_commit_batch_dev): list of CUDA tensors of Montgomery Fr -> (count, 48) uint8 CUDA tensor"""
        import torch

        count = len(coeff_tensors)
        dev = coeff_tensors[0].device if count else torch.device("cuda", torch.cuda.current_device())
        if out is None:
            out = torch.empty((max(count, 1), 48), dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * max(count, 1))(*[t.data_ptr() for t in coeff_tensors])
        lens = (C.c_size_t * max(count, 1))(*[t.numel() * t.element_size() // 32 for t in coeff_tensors])
        with torch.cuda.device(dev):
            srs._lib.check(srs._lib.kzg_commit_batch_dev(srs._h, out.data_ptr(), ptrs, lens, count,
                                                         torch.cuda.current_stream().cuda_stream), "aleo_b200_kzg_commit_batch_dev")
        return out[:count]

    @staticmethod
    def commit_hiding_dev(srs_beta: ResidentSRS, srs_beta_gamma: ResidentSRS, coeffs_t, random_coeffs_t, out=None):
        """KZG10::commit with a hiding bound: msm(powers_of_beta_g, coeffs) + msm(powers_of_beta_times_gamma_g, random)"""
        import torch

        if out is None:
            out = torch.empty(48, dtype=torch.uint8, device=coeffs_t.device)
        with torch.cuda.device(coeffs_t.device):
            srs_beta._lib.check(srs_beta._lib.kzg_commit_hiding_dev(
                srs_beta._h, srs_beta_gamma._h, out.data_ptr(), coeffs_t.data_ptr(), coeffs_t.numel() * coeffs_t.element_size() // 32,
                random_coeffs_t.data_ptr(), random_coeffs_t.numel() * random_coeffs_t.element_size() // 32,
                torch.cuda.current_stream().cuda_stream), "aleo_b200_kzg_commit_hiding_dev")
        return out

    @staticmethod
    def commit_lagrange_dev(srs_lagrange: ResidentSRS, evaluations_t, n_evals: int, out=None):
        """KZG10::commit_lagrange: the same msm over a handle built from lagrange_basis_at_beta_g, fed with the
        evaluations over the domain instead of coefficients (lets the prover skip an ifft)"""
        return KZG10.commit_dev(srs_lagrange, evaluations_t, n_evals, out)

    @staticmethod
    def open_dev(srs: ResidentSRS, coeffs_t, point, out=None):
        """non-hiding KZG10::open at `point` (canonical int): witness polynomial (p(x) - p(z)) / (x - z) built on the
        device, then its commitment against the resident powers -> 48-byte compressed G1 (CUDA tensor)"""
        import torch

        from .poly import _fr_host
        if out is None:
            out = torch.empty(48, dtype=torch.uint8, device=coeffs_t.device)
        n = coeffs_t.numel() * coeffs_t.element_size() // 32
        with torch.cuda.device(coeffs_t.device):
            srs._lib.check(srs._lib.kzg_open_dev(srs._h, out.data_ptr(), coeffs_t.data_ptr(), n, _fr_host(point),
                                                 torch.cuda.current_stream().cuda_stream), "aleo_b200_kzg_open_dev")
        return out

    @staticmethod
    def open_combinations_dev(srs: ResidentSRS, polys, lc_rows, points, out=None):
        """m openings in one call (SonicKZG10::open_combinations / batch_open shape): opening k is of the linear
        combination sum_i lc_rows[k][i] * polys[i] at points[k]; coefficients and points are canonical ints, `polys` a
        list of CUDA tensors of Montgomery Fr.  -> (m, 48) uint8 CUDA tensor of compressed witness commitments."""
        import torch

        from .poly import _fr_host
        m, npoly = len(lc_rows), len(polys)
        if m != len(points) or any(len(r) != npoly for r in lc_rows):
            raise ValueError("lc_rows must be m rows of len(polys) coefficients, points m values")
        dev = polys[0].device
        if out is None:
            out = torch.empty((max(m, 1), 48), dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * max(npoly, 1))(*[t.data_ptr() for t in polys])
        lens = (C.c_size_t * max(npoly, 1))(*[t.numel() * t.element_size() // 32 for t in polys])
        lc = b"".join(_fr_host(a) for row in lc_rows for a in row)
        zs = b"".join(_fr_host(z) for z in points)
        with torch.cuda.device(dev):
            srs._lib.check(srs._lib.kzg_open_combinations_dev(srs._h, out.data_ptr(), ptrs, lens, npoly, lc, zs, m,
                                                              torch.cuda.current_stream().cuda_stream),
                           "aleo_b200_kzg_open_combinations_dev")
        return out[:m]
