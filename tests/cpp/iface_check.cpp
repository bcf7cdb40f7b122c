// tests/cpp/iface_check.cpp -- drives the two HIGH-LEVEL C++ interfaces of include/tlib/ttv.h on the GPU:
//   (1) auto C = A(q) * b;                       operator*            reference include/tlib/ttv.h:122-127
//   (2) auto C = ttv(q, A, b, ep, sp, fp);       tensor-level         reference include/tlib/ttv.h:99-114
// as the reference's example/interface1.cpp:43 and interface2.cpp:43-44 use them, plus this repo's device-resident forms
// (tensor::keep_on_device, device_tensor).  Test infrastructure: reads one case written by tests/test_cpp_interfaces_gpu.py
// (element type, order, q, shape, layout, A, b), writes every result back; the Python side compares with the oracle.
//
//   iface_check <cases.bin> <out.bin>
// cases.bin: any number of cases back to back; a case = int64 dtype, p, q, shape[p], layout[p]; then A (N elements),
//            b (n_q elements), raw.
// out.bin  : per case 6 result arrays of N / n_q elements each, in the order of run() below.
#include <tlib/ttv.h>

#include <complex>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

using namespace tlib::ttv;

static void read_exact(std::FILE* f, void* dst, std::size_t bytes)
{
  if (bytes && std::fread(dst, 1, bytes, f) != bytes) throw std::runtime_error("short read");
}

template<class T>
static void put(std::FILE* f, tensor<T> const& c, std::size_t want)
{
  if (c.data().size() != want) throw std::runtime_error("result has " + std::to_string(c.data().size()) + " elements, expected " + std::to_string(want));
  if (std::fwrite(c.data().data(), sizeof(T), want, f) != want) throw std::runtime_error("short write");
}

template<class T>
static int run(std::FILE* in, std::FILE* out, std::size_t p, std::size_t q, std::vector<std::size_t> const& n, std::vector<std::size_t> const& pi)
{
  tensor<T> A(n, pi);
  tensor<T> B({n.at(q - 1), 1});                                   // a vector as the examples build it: shape {n_q, 1}
  read_exact(in, A.data().data(), A.data().size() * sizeof(T));
  read_exact(in, B.data().data(), B.data().size() * sizeof(T));
  std::size_t const n_out = A.data().size() / n.at(q - 1);
  (void)p;

  // [0] interface 1 on plain host tensors
  { auto C = A(q) * B; put(out, C, n_out); }
  // [1] interface 2 with another policy triple (all of them select the same GPU path)
  { auto C = ttv(q, A, B, execution_policy::par_loop, slicing_policy::slice, fusion_policy::all); put(out, C, n_out); }
  // [2], [3] a host tensor that keeps its copy in HBM: the first product uploads, the second one must not
  A.keep_on_device();
  { auto C = A(q) * B; put(out, C, n_out); }
  if (!ttv_b200_resident_valid(A.device_twin())) throw std::runtime_error("the tensor did not keep its device copy");
  { auto C = ttv(q, A, B, execution_policy::seq, slicing_policy::subtensor, fusion_policy::none); put(out, C, n_out); }
  if (!ttv_b200_resident_valid(A.device_twin())) throw std::runtime_error("the device copy was dropped by a product");
  // [4] operands and result on the device
  {
    device_tensor<T> dA(A), dB(B);
    auto dC = dA(q) * dB;
    put(out, dC.to_host(), n_out);
  }
  // [5] the host data changes through the container's accessors: the device copy must be refreshed
  *A.begin() = *A.begin() + T(1);
  if (ttv_b200_resident_valid(A.device_twin())) throw std::runtime_error("mutable access did not invalidate the device copy");
  { auto C = A(q) * B; put(out, C, n_out); }
  return 0;
}

int main(int argc, char** argv)
{
  if (argc != 3) { std::fprintf(stderr, "usage: iface_check case.bin out.bin\n"); return 2; }
  try {
    std::FILE* in = std::fopen(argv[1], "rb");
    std::FILE* out = std::fopen(argv[2], "wb");
    if (!in || !out) throw std::runtime_error("cannot open files");
    int rc = 0;
    for (;;) {
      std::int64_t head[3];
      std::size_t const got = std::fread(head, 1, sizeof head, in);
      if (got == 0) break;                                           // end of the cases
      if (got != sizeof head) throw std::runtime_error("short read");
      std::size_t const p = static_cast<std::size_t>(head[1]), q = static_cast<std::size_t>(head[2]);
      std::vector<std::int64_t> raw(2 * p);
      read_exact(in, raw.data(), raw.size() * sizeof(std::int64_t));
      std::vector<std::size_t> n(raw.begin(), raw.begin() + p), pi(raw.begin() + p, raw.end());
      switch (head[0]) {
        case 0: rc |= run<float>(in, out, p, q, n, pi); break;
        case 1: rc |= run<double>(in, out, p, q, n, pi); break;
        case 2: rc |= run<std::complex<float>>(in, out, p, q, n, pi); break;
        case 3: rc |= run<std::complex<double>>(in, out, p, q, n, pi); break;
        case 4: rc |= run<std::int32_t>(in, out, p, q, n, pi); break;
        case 5: rc |= run<std::int64_t>(in, out, p, q, n, pi); break;
        default: throw std::runtime_error("unknown element type");
      }
    }
    std::fclose(in);
    std::fclose(out);
    return rc;
  } catch (std::exception const& e) {
    std::fprintf(stderr, "iface_check: %s\n", e.what());
    return 1;
  }
}
