// instantiates the DAE kernel (dae.cuh) for model 'calvetti' (index-1 DAE with time-triggered switches, reference models/model_calvetti_py).
#include "../dae_launch.cuh"
#include "../models/calvetti.cuh"
namespace amici_b200 {
ModelEntry entry_calvetti() { return make_dae_entry<models::calvetti>(); } }
