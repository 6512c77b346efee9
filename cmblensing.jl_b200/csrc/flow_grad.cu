// Pullback through LenseFlow (negδvelocityᴴ, src/lenseflow.jl:176-214) — implemented in a later step.
#include "flow.cuh"
namespace cmbl {
template <class T> void flow_grad(FlowT<T>&, int, const T*, const C2<T>*, C2<T>*, C2<T>*, bool, cmblStream_t) {
    throw Error("cmbl_lenseflow_grad: not implemented yet");
}
template void flow_grad<float>(FlowT<float>&, int, const float*, const C2<float>*, C2<float>*, C2<float>*, bool, cmblStream_t);
template void flow_grad<double>(FlowT<double>&, int, const double*, const C2<double>*, C2<double>*, C2<double>*, bool, cmblStream_t);
}
