// placeholder until the tcgen05 persistent sweep lands
#include "common.cuh"
namespace ds2 {
struct SeqArgs;
int rnn_sweep_fwd_tc(int, const SeqArgs&, void*, size_t, cudaStream_t) { return 1; }
int rnn_sweep_bwd_tc(int, const SeqArgs&, void*, size_t, cudaStream_t) { return 1; }
size_t rnn_sweep_tc_workspace_bytes(int, int, int, int, int) { return 0; }
}
