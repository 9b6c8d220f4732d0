// test_mirror.cpp -- exercises the C++ host-side mirror (include/aleo_b200.hpp) the way snarkVM's own
// msm / fft tests read: msm against scalar-times-point identities, fft round trips, domain rules.
//   g++ -std=c++17 -Iinclude tests/cpp/test_mirror.cpp -Laleo_b200 -laleo_b200 -Wl,-rpath,$PWD/aleo_b200 -o build/test_mirror
// Without arguments it only checks host logic (layouts, EvaluationDomain::new); with "gpu" it calls the device.
#include <cstdio>
#include <cstring>
#include <vector>

#include "aleo_b200.hpp"

using namespace aleo_b200;

#define REQUIRE(cond)                                                        \
  do {                                                                       \
    if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  // EvaluationDomain::new rules (upstream: next power of two, None above the two-adicity)
  REQUIRE(EvaluationDomain::new_(1)->size == 1);
  REQUIRE(EvaluationDomain::new_(3)->size == 4);
  REQUIRE(EvaluationDomain::new_(1u << 20)->log_size_of_group == 20);
  REQUIRE(EvaluationDomain::new_((size_t(1) << 20) + 1)->log_size_of_group == 21);
  REQUIRE(!EvaluationDomain::new_((size_t(1) << 47) + 1).has_value());
  if (argc < 2 || std::strcmp(argv[1], "gpu") != 0) {
    // no device: the mirror must throw, not fall back
    bool threw = false;
    try {
      std::vector<Fr> v(8);
      EvaluationDomain::new_(8)->fft_in_place(v);
    } catch (const std::runtime_error&) { threw = true; }
    if (aleo_b200_device_count() == 0) REQUIRE(threw);
    std::printf("host logic ok\n");
    return 0;
  }
  // msm of n copies of the identity / zero scalars is the identity (0, 1, 0)
  std::vector<G1Affine> bases(5);
  std::vector<BigInteger256> scalars(7);
  std::memset(bases.data(), 0, bases.size() * sizeof(G1Affine));
  std::memset(scalars.data(), 0, scalars.size() * sizeof(BigInteger256));
  for (auto& b : bases) b.infinity = true;
  G1Projective r = VariableBase::msm(bases, scalars);     // ragged: 5 bases, 7 scalars -> zipped
  uint64_t zacc = 0;
  for (int i = 0; i < 6; i++) zacc |= r.z.v[i];
  REQUIRE(zacc == 0);
  // fft then ifft is the identity; fft_in_place resizes a 5-vector to the domain of 8
  auto dom = *EvaluationDomain::new_(5);
  std::vector<Fr> v(5);
  for (size_t i = 0; i < v.size(); i++) v[i] = Fr{{i + 1, 0, 0, 0}};   // small residues are valid field elements
  std::vector<Fr> w = v;
  dom.fft_in_place(w);
  REQUIRE(w.size() == 8);
  dom.ifft_in_place(w);
  for (size_t i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) REQUIRE(w[i].v[k] == (i < 5 ? v[i].v[k] : 0));
  std::vector<Fr> c = dom.coset_ifft(dom.coset_fft(v));
  for (size_t i = 0; i < 5; i++) REQUIRE(c[i].v[0] == v[i].v[0]);
  // FFTOrder: IO followed by OI of the inverse is the identity (the bit-reversed intermediate is consumed as it is)
  {
    std::vector<Fr> u = v;
    dom.out_order_fft_in_place(u);
    dom.ifft_helper_in_place(u, EvaluationDomain::FFTOrder::OI);
    for (size_t i = 0; i < 8; i++) REQUIRE(u[i].v[0] == (i < 5 ? v[i].v[0] : 0));
  }
  // PolyMultiplier: one polynomial and no evaluations gives the polynomial back, zero padded to the domain
  {
    PolyMultiplier pm;
    pm.add_polynomial(v);
    std::vector<Fr> p1 = pm.multiply(dom);
    REQUIRE(p1.size() == 8);
    for (size_t i = 0; i < 8; i++)
      for (int k = 0; k < 4; k++) REQUIRE(p1[i].v[k] == (i < 5 ? v[i].v[k] : 0));
    // (p * p) evaluated on the domain == pointwise square of fft(p): multiply, then transform both ways
    auto dom16 = *EvaluationDomain::new_(16);
    PolyMultiplier sq;
    sq.add_polynomial(v);
    sq.add_polynomial(v);
    std::vector<Fr> pp = sq.multiply(dom16);
    PolyMultiplier viaev;
    std::vector<Fr> e = dom16.fft(v);
    viaev.add_polynomial(v);
    viaev.add_evaluation(e);
    std::vector<Fr> pe = viaev.multiply(dom16);
    for (size_t i = 0; i < 16; i++)
      for (int k = 0; k < 4; k++) REQUIRE(pp[i].v[k] == pe[i].v[k]);
    for (size_t i = 9; i < 16; i++) REQUIRE((pp[i].v[0] | pp[i].v[1] | pp[i].v[2] | pp[i].v[3]) == 0);   // degree 8
  }
  // KZG10::commit against resident powers: all-identity powers commit to the identity (flag bit 382 set, x = 0),
  // and a handle serves any prefix length
  {
    std::vector<G1Affine> powers(300);
    std::memset(powers.data(), 0, powers.size() * sizeof(G1Affine));
    for (auto& b : powers) b.infinity = true;
    ResidentPowers rp(powers);
    REQUIRE(rp.size() == 300);
    std::vector<Fr> poly(123, Fr{{5, 0, 0, 0}});
    KZGCommitment cm = KZG10::commit(rp, poly);
    REQUIRE(cm.bytes[47] == 0x40);
    for (int i = 0; i < 47; i++) REQUIRE(cm.bytes[i] == 0);
    G1Projective z = rp.msm(std::vector<BigInteger256>(7, BigInteger256{{3, 0, 0, 0}}));
    uint64_t acc = 0;
    for (int i = 0; i < 6; i++) acc |= z.z.v[i];
    REQUIRE(acc == 0);
  }
  std::printf("gpu mirror ok\n");
  return 0;
}
