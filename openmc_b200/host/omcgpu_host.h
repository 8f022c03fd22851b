// omcgpu_host.h -- libopenmc-side hooks of the B200 transport path (see omcgpu_host.cpp).
//
// Four call sites of libopenmc change (apply_hooks.py makes exactly these edits; INTEGRATION.md
// shows them as a diff):
//
//   src/simulation.cpp:99-105   openmc_simulation_init():  `if (settings::event_based) init_event_queues(...)`
//                               -> `if (gpu::enabled()) gpu::simulation_init(); else if (...)`
//   src/simulation.cpp:255-259  openmc_next_batch():        `if (settings::event_based) transport_event_based();
//                               else transport_history_based();`
//                               -> `if (gpu::enabled()) gpu::transport_generation(); else if (...)`
//   src/simulation.cpp:546-552  finalize_generation():      `sort_fission_bank(); synchronize_bank();`
//                               -> `if (!gpu::bank_synchronized()) { ... }` (the device sorted and combed the bank)
//   src/finalize.cpp:57-59      free_memory():              `gpu::simulation_finalize();` next to free_event_queues()
//
// Everything else of libopenmc -- XML input, data loading, source sampling, synchronize_bank, k statistics,
// tally accumulation, statepoint / summary / tallies.out, the openmc_* C API -- runs unchanged.
#ifndef OMCGPU_HOST_H
#define OMCGPU_HOST_H

struct omcgpu_model;

namespace openmc {
namespace gpu {

//! OPENMC_GPU set to anything but 0: evaluated next to settings::event_based
bool enabled();
//! model:: / data:: / settings:: -> the flattened POD model; throws std::runtime_error naming the
//! feature when the model needs something the device path refuses
const omcgpu_model* flatten_model();
void simulation_init();
void transport_generation();
//! true: sort_fission_bank() / synchronize_bank() of this generation were done on the device
bool bank_synchronized();
// shannon_entropy (src/eigenvalue.cpp:579-607) over the fission bank that stayed on the device
void shannon_entropy();
void simulation_finalize();

} // namespace gpu
} // namespace openmc

// C entry points next to the openmc_* API
extern "C" const omcgpu_model* openmc_gpu_flatten_model(void);
#endif
