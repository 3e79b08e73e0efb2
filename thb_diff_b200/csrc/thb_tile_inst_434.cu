// Tile-kernel instantiations for tensor-product shape (4,3,4) = degrees (3,2,3); generated list, see thb_tile.cu
#include "thb_net_impl.cuh"
THB_DEFINE_SHAPE(4, 3, 4)
