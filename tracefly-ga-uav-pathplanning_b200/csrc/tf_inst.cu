// tf_inst.cu — one translation unit per (kernel family, NW): compiled several times by build.py with
//   -DTF_INST_KIND=<0 fitness LSU | 1 generation LSU | 2 fitness stream | 3 generation stream> -DTF_INST_NW=<n>
// so that the template instantiations build in parallel.
#ifndef TF_INST_KIND
#error "compile with -DTF_INST_KIND=... -DTF_INST_NW=..."
#endif
#define TF_CAT2(a, b) a##b
#define TF_CAT(a, b) TF_CAT2(a, b)

#if TF_INST_KIND == 0
#include "tf_fitness_kernels.cuh"
namespace tf {
int TF_CAT(fitness_lsu_nw, TF_INST_NW)(const EnvDev& e, const tf_params& prm, const float* pop, const int32_t* lens, int P,
                                       int W, float* fit, uint32_t* coll, cudaStream_t st) {
    return dispatch_sem<TF_INST_NW>(e, prm, pop, lens, P, W, fit, coll, st);
}
}  // namespace tf
#elif TF_INST_KIND == 1
#include "tf_generation_kernels.cuh"
namespace tf {
int TF_CAT(generation_lsu_nw, TF_INST_NW)(const GenArgs& a) { return gen_sem<TF_INST_NW>(a); }
}  // namespace tf
#elif TF_INST_KIND == 2
#include "tf_stream_dev.cuh"
namespace tf {
#if TF_INST_NW >= 4
int TF_CAT(fitness_stream_nw, TF_INST_NW)(const StreamArgs& a) { return stream_long_rows<TF_INST_NW, false>(a); }
#else
int TF_CAT(fitness_stream_nw, TF_INST_NW)(const StreamArgs& a) { return stream_sem<TF_INST_NW, false>(a); }
#endif
}  // namespace tf
#elif TF_INST_KIND == 3
#include "tf_stream_dev.cuh"
namespace tf {
#if TF_INST_NW >= 4
int TF_CAT(generation_stream_nw, TF_INST_NW)(const StreamArgs& a) { return stream_long_rows<TF_INST_NW, true>(a); }
#else
int TF_CAT(generation_stream_nw, TF_INST_NW)(const StreamArgs& a) { return stream_sem<TF_INST_NW, true>(a); }
#endif
}  // namespace tf
#endif
