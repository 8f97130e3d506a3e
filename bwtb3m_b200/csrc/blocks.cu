#include "engine.h"
namespace b3m {
void Engine::build_blocks(PhaseTimer &, uint32_t *) { throw Error("numblocks > 1 not implemented yet"); }
}
