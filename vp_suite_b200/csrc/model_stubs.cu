// Temporary: rollouts not yet implemented fail loudly.
#include "cells.h"
#include "model.h"
namespace vpk {
Model* make_predrnn(const vpk_model_desc&) { VPK_THROW(3, "predrnn-pp rollout not built yet"); }
Model* make_phydnet(const vpk_model_desc&, bool) { VPK_THROW(3, "phy rollout not built yet"); }
Cell* make_phycell_cell(int, int, int, int, int, int, int, const float*, const float*, const float*, const float*,
                        const float*, const float*, const float*, const float*) {
  VPK_THROW(3, "PhyCell cell not built yet");
}
}  // namespace vpk
