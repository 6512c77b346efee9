// Helpers shared by the extern "C" translation units: error capture, handle definitions, dtype dispatch.
#pragma once
#include "plan.cuh"
#include "../../include/cmbl_b200.h"

namespace cmbl {
void set_last_error(const std::string& s);
struct FlowBase;   // flow.cuh
struct CgBase;     // cg.cuh
}

struct cmbl_plan { std::unique_ptr<cmbl::PlanBase> p; };
#define CMBL_FLOW_STRUCT struct cmbl_flow { std::unique_ptr<cmbl::FlowBase> f; cmbl_plan* plan; }

#define CMBL_API_BEGIN try {
#define CMBL_API_END                                                                          \
    return CMBL_OK;                                                                           \
    } catch (const ::cmbl::Error& e) { ::cmbl::set_last_error(e.what());                      \
        return (std::string(e.what()).find("cuda") != std::string::npos) ? CMBL_ERR_CUDA : CMBL_ERR_INVALID; } \
    catch (const std::exception& e) { ::cmbl::set_last_error(e.what()); return CMBL_ERR_INVALID; }      \
    catch (...) { ::cmbl::set_last_error("unknown error"); return CMBL_ERR_INVALID; }

// run `expr` with PT bound to PlanT<float> or PlanT<double>
#define CMBL_DISPATCH(planbase, ...)                                                           \
    do { if ((planbase)->dtype == 0) { typedef float T; auto& P = *static_cast<::cmbl::PlanT<float>*>(planbase); (void)P; __VA_ARGS__; } \
         else { typedef double T; auto& P = *static_cast<::cmbl::PlanT<double>*>(planbase); (void)P; __VA_ARGS__; } } while (0)

static inline cmblStream_t as_stream(void* s) { return reinterpret_cast<cmblStream_t>(s); }
