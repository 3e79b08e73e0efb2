// Tile-kernel instantiations for tensor-product shape (3,3,3) = degrees (2,2,2); generated list, see thb_tile.cu
#include "thb_net_impl.cuh"
THB_DEFINE_SHAPE(3, 3, 3)
