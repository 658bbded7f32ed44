// GENERATED: instantiates the block-solver kernel for the mass-action network 'massaction'.
#include "../block_launch.cuh"
#include "../models/massaction.cuh"
namespace amici_b200 {
ModelEntry entry_massaction() { return make_block_entry<models::massaction>(); } }
