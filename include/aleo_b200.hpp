// aleo_b200.hpp -- C++ host-side mirror of the snarkVM 0.14.5 interface for the hot path, above the C ABI
// (include/aleo_b200.h).  The reference's host language is Rust and this image has no Rust toolchain, so the
// compiled-language mirror is C++: same names, argument meaning and error behaviour as
//   snarkvm_algorithms::msm::VariableBase::msm                      (src/msm/variable_base/mod.rs)
//   snarkvm_algorithms::fft::EvaluationDomain::{new, fft, ifft, coset_fft, coset_ifft, *_in_place}
//                                                                    (src/fft/domain.rs)
// reached from the reference at rust/src/program/execute.rs:74,177.  Struct layouts are the Rust memory images
// the FFI receives; the static_asserts below are the layout contract the Rust shim relies on.
// Errors: the Rust shim panics on a non-zero return code (no CPU fallback); here that is a std::runtime_error.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "aleo_b200.h"

namespace aleo_b200 {

struct BigInteger256 { uint64_t v[4]; };                 // canonical little-endian
struct Fr { uint64_t v[4]; };                            // Fp256<FrParameters>: Montgomery residue
struct Fq { uint64_t v[6]; };                            // Fp384<FqParameters>: Montgomery residue
struct G1Affine { Fq x, y; bool infinity; };             // short_weierstrass_jacobian::Affine<Bls12_377G1Parameters>
struct G1Projective { Fq x, y, z; };                     // Jacobian; identity: z == 0

static_assert(sizeof(BigInteger256) == 32 && sizeof(Fr) == 32 && sizeof(Fq) == 48, "field element images");
static_assert(sizeof(G1Affine) == 104 && offsetof(G1Affine, y) == 48 && offsetof(G1Affine, infinity) == 96,
              "size_of::<G1Affine>() == 104: x | y | infinity | 7 bytes padding");
static_assert(sizeof(G1Projective) == 144, "G1Projective image");

inline void check(int rc, const char* what) {
  if (rc != ALEO_B200_OK)
    throw std::runtime_error(std::string(what) + ": " + aleo_b200_strerror(rc) + " [" + aleo_b200_last_cuda_error() + "]");
}

struct VariableBase {
  // VariableBase::msm(bases, scalars) -> G::Projective; the two slices are zipped (shorter length wins)
  static G1Projective msm(const G1Affine* bases, size_t n_bases, const BigInteger256* scalars, size_t n_scalars) {
    G1Projective out;
    const size_t n = n_bases < n_scalars ? n_bases : n_scalars;
    check(aleo_b200_msm_g1(&out, bases, n, scalars, sizeof(G1Affine)), "VariableBase::msm");
    return out;
  }
  static G1Projective msm(const std::vector<G1Affine>& bases, const std::vector<BigInteger256>& scalars) {
    return msm(bases.data(), bases.size(), scalars.data(), scalars.size());
  }
};

class EvaluationDomain {
 public:
  static constexpr uint32_t TWO_ADICITY = 47;
  uint64_t size = 1;
  uint32_t log_size_of_group = 0;

  // EvaluationDomain::new(num_coeffs): next power of two; None when log2(size) > two-adicity
  static std::optional<EvaluationDomain> new_(size_t num_coeffs) {
    EvaluationDomain d;
    while (d.size < num_coeffs) {
      d.size <<= 1;
      d.log_size_of_group++;
      if (d.log_size_of_group > TWO_ADICITY) return std::nullopt;
    }
    return d;
  }
  // *_in_place first resize the vector to the domain size with zeros (a longer input is truncated)
  void fft_in_place(std::vector<Fr>& v) const { run(v, ALEO_B200_NTT_FORWARD, ALEO_B200_NTT_STANDARD); }
  void ifft_in_place(std::vector<Fr>& v) const { run(v, ALEO_B200_NTT_INVERSE, ALEO_B200_NTT_STANDARD); }
  void coset_fft_in_place(std::vector<Fr>& v) const { run(v, ALEO_B200_NTT_FORWARD, ALEO_B200_NTT_COSET); }
  void coset_ifft_in_place(std::vector<Fr>& v) const { run(v, ALEO_B200_NTT_INVERSE, ALEO_B200_NTT_COSET); }
  std::vector<Fr> fft(const std::vector<Fr>& c) const { auto v = c; fft_in_place(v); return v; }
  std::vector<Fr> ifft(const std::vector<Fr>& e) const { auto v = e; ifft_in_place(v); return v; }
  std::vector<Fr> coset_fft(const std::vector<Fr>& c) const { auto v = c; coset_fft_in_place(v); return v; }
  std::vector<Fr> coset_ifft(const std::vector<Fr>& e) const { auto v = e; coset_ifft_in_place(v); return v; }

  // FFTOrder variants (fft_helper_in_place_with_pc & co.): II natural in / natural out, IO bit-reversed out, OI
  // bit-reversed in.  The FFTPrecomputation argument of upstream is dropped: twiddle tables are cached in the library.
  enum class FFTOrder : int { II = ALEO_B200_NTT_ORDER_II, IO = ALEO_B200_NTT_ORDER_IO, OI = ALEO_B200_NTT_ORDER_OI };
  void fft_helper_in_place(std::vector<Fr>& v, FFTOrder order) const { run_ordered(v, ALEO_B200_NTT_FORWARD, ALEO_B200_NTT_STANDARD, order); }
  void ifft_helper_in_place(std::vector<Fr>& v, FFTOrder order) const { run_ordered(v, ALEO_B200_NTT_INVERSE, ALEO_B200_NTT_STANDARD, order); }
  void out_order_fft_in_place(std::vector<Fr>& v) const { fft_helper_in_place(v, FFTOrder::IO); }

 private:
  void run_ordered(std::vector<Fr>& v, int direction, int kind, FFTOrder order) const {
    v.resize(size, Fr{{0, 0, 0, 0}});
    check(aleo_b200_ntt_fr_ordered(v.data(), log_size_of_group, direction, kind, (int)order), "EvaluationDomain::fft_helper");
  }
  void run(std::vector<Fr>& v, int direction, int kind) const {
    v.resize(size, Fr{{0, 0, 0, 0}});
    check(aleo_b200_ntt_fr(v.data(), log_size_of_group, direction, kind), "EvaluationDomain::fft");
  }
};

// PolyMultiplier (src/fft/polynomial/multiplier.rs): collect polynomials and evaluation vectors, multiply() returns the
// coefficients of ifft(prod fft(p_i) * prod e_j) over `domain` -- every operand crosses PCIe once (aleo_b200_polymul).
class PolyMultiplier {
 public:
  void add_polynomial(const std::vector<Fr>& p) { polys_.push_back(&p); }
  void add_evaluation(const std::vector<Fr>& e) { evals_.push_back(&e); }
  std::vector<Fr> multiply(const EvaluationDomain& domain) const {
    std::vector<const void*> pp, ep;
    std::vector<size_t> pl, el;
    for (auto* p : polys_) { pp.push_back(p->data()); pl.push_back(p->size()); }
    for (auto* e : evals_) { ep.push_back(e->data()); el.push_back(e->size()); }
    std::vector<Fr> out(domain.size);
    check(aleo_b200_polymul(out.data(), pp.size(), pp.data(), pl.data(), ep.size(), ep.data(), el.data(), domain.log_size_of_group),
          "PolyMultiplier::multiply");
    return out;
  }

 private:
  std::vector<const std::vector<Fr>*> polys_, evals_;
};

// Commitment as it appears in a proof: compressed G1 (x little-endian | bit 383: larger y | bit 382: infinity)
struct KZGCommitment { uint8_t bytes[48]; };
static_assert(sizeof(KZGCommitment) == 48, "compressed G1");

// `Powers` whose powers_of_beta_g (or lagrange_basis_at_beta_g) stay resident on the device (aleo_b200_srs_*):
// uploaded and expanded once, then every commitment only moves its coefficients.  Move-only.
class ResidentPowers {
 public:
  explicit ResidentPowers(const std::vector<G1Affine>& powers_of_beta_g) {
    check(aleo_b200_srs_create(&h_, powers_of_beta_g.data(), powers_of_beta_g.size(), sizeof(G1Affine)), "srs_create");
  }
  ResidentPowers(const ResidentPowers&) = delete;
  ResidentPowers& operator=(const ResidentPowers&) = delete;
  ResidentPowers(ResidentPowers&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ~ResidentPowers() {
    if (h_) aleo_b200_srs_destroy(h_);
  }
  const void* handle() const { return h_; }
  size_t size() const {
    size_t n = 0;
    check(aleo_b200_srs_info(h_, &n, nullptr, nullptr, nullptr), "srs_info");
    return n;
  }
  // VariableBase::msm over a prefix of the resident bases
  G1Projective msm(const std::vector<BigInteger256>& scalars) const {
    G1Projective out;
    check(aleo_b200_srs_msm(h_, &out, scalars.data(), scalars.size()), "srs_msm");
    return out;
  }

 private:
  void* h_ = nullptr;
};

struct KZG10 {
  // KZG10::commit(powers, polynomial, hiding_bound = None): coefficients as the polynomial holds them (Montgomery Fr)
  static KZGCommitment commit(const ResidentPowers& powers, const std::vector<Fr>& coeffs) {
    KZGCommitment c;
    check(aleo_b200_kzg_commit(powers.handle(), c.bytes, coeffs.data(), coeffs.size()), "KZG10::commit");
    return c;
  }
  // KZG10::commit_lagrange(lagrange_basis, evaluations): `lagrange_basis` = a handle built from lagrange_basis_at_beta_g
  static KZGCommitment commit_lagrange(const ResidentPowers& lagrange_basis, const std::vector<Fr>& evaluations) {
    return commit(lagrange_basis, evaluations);
  }
};

}  // namespace aleo_b200
