// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// extern "C" doorway onto the UNMODIFIED reference CFAR source.  The reference
// file is compiled from where it lies (REF_CFAR_CPP is set by oracle/Makefile to
// /root/reference/bruce_slam/src/bruce_slam/cpp/cfar.cpp) against the container
// shims in oracle/shim/; nothing of it is copied into this repository.  The
// result, oracle/_ref/libcfar_ref.so, is git-ignored and travels to the GPU box
// as a prebuilt file.
//
// What pybind11 does at the reference boundary is reproduced here by hand:
// the numpy image (row-major, any dtype) is converted into a fresh float32
// column-major MatrixXf (cfar.cpp:10 `const MatrixXf &img`), and the returned
// column-major matrices are copied out (we hand them back row-major; the
// Python wrapper restores the F-order the pybind module returns).
#include REF_CFAR_CPP

#include <cstring>

extern "C" {

// alg: 0 CA, 1 SOCA, 2 GOCA, 3 OS   (cfar.cpp:10,30,53,76 / :98,120,145,170)
// thr_out == NULL -> plain variant, else the "*2" variant.
int ref_cfar(int alg, const float *img, int R, int B, int train_hs, int guard_hs, int k,
             double tau, uint8_t *mask_out, float *thr_out) {
  MatrixXf m(R, B);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < B; ++c) m(r, c) = img[(size_t)r * B + c];
  MatrixXb ret;
  MatrixXf ret2;
  if (!thr_out) {
    switch (alg) {
      case 0: ret = ca(m, train_hs, guard_hs, tau); break;
      case 1: ret = soca(m, train_hs, guard_hs, tau); break;
      case 2: ret = goca(m, train_hs, guard_hs, tau); break;
      case 3: ret = os(m, train_hs, guard_hs, k, tau); break;
      default: return -1;
    }
  } else {
    std::pair<MatrixXb, MatrixXf> p;
    switch (alg) {
      case 0: p = ca2(m, train_hs, guard_hs, tau); break;
      case 1: p = soca2(m, train_hs, guard_hs, tau); break;
      case 2: p = goca2(m, train_hs, guard_hs, tau); break;
      case 3: p = os2(m, train_hs, guard_hs, k, tau); break;
      default: return -1;
    }
    ret = p.first;
    ret2 = p.second;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < B; ++c) thr_out[(size_t)r * B + c] = ret2(r, c);
  }
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < B; ++c) mask_out[(size_t)r * B + c] = ret(r, c);
  return 0;
}

}  // extern "C"
