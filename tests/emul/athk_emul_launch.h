// TEST INFRASTRUCTURE (tests/emul): the launchers of athk_backend.h as plain host loops (included
// inside namespace athk by athk_backend.h when ATHK_EMULATE is defined).
template <class F>
static inline cudaError_t launch_links(cudaStream_t, int nlinks, long long maxcount, const F &f) {
  for (int l = 0; l < nlinks; ++l)
    for (long long e = 0; e < maxcount; ++e) f(l, e);
  return 0;
}
template <class F>
static inline cudaError_t launch_points(cudaStream_t, int nb, const Range3 &r, const F &f,
                                        int = 256) {
  for (int b = 0; b < nb; ++b)
    for (int k = r.kl; k <= r.ku; ++k)
      for (int j = r.jl; j <= r.ju; ++j)
        for (int i = r.il; i <= r.iu; ++i) f(b, k, j, i);
  return 0;
}
template <int ORDER, bool MHD, int RS, int AR, class F>
static inline cudaError_t launch_sweep_x1(cudaStream_t st, int nb, const Range3 &r, const F &f) {
  return launch_points(st, nb, r, f);
}
template <int MINB, class F>
static inline cudaError_t launch_points_occ(cudaStream_t st, int nb, const Range3 &r, const F &f) {
  return launch_points(st, nb, r, f);
}
template <int MINB, class F>
static inline cudaError_t launch_points_kcol(cudaStream_t, int nb, const Range3 &r, int kt, const F &f) {
  if (kt < 1) kt = 1;
  for (int b = 0; b < nb; ++b)
    for (int k0 = r.kl; k0 <= r.ku; k0 += kt)
      for (int j = r.jl; j <= r.ju; ++j)
        for (int i = r.il; i <= r.iu; ++i) f.col(b, k0, (k0 + kt - 1 < r.ku) ? k0 + kt - 1 : r.ku, j, i);
  return 0;
}
template <int MINB, class F>
static inline cudaError_t launch_points_min(cudaStream_t st, int nb, const Range3 &r, const F &f, double *) {
  return launch_points(st, nb, r, f);   // operator() folds cell by cell
}
template <class F>
static inline cudaError_t launch_shell(cudaStream_t st, int nb, const Range3 &outer, const Range3 &, const F &f) {
  return launch_points(st, nb, outer, f);   // the functors skip the points of the inner box themselves
}
