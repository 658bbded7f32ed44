// GENERATED: instantiates the block-solver kernel for the mass-action network 'massaction_small'.
#include "../block_launch.cuh"
#include "../models/massaction_small.cuh"
namespace amici_b200 {
ModelEntry entry_massaction_small() { return make_block_entry<models::massaction_small>(); } }
