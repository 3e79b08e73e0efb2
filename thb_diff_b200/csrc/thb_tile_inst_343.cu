// Tile-kernel instantiations for tensor-product shape (3,4,3) = degrees (2,3,2); generated list, see thb_tile.cu
#include "thb_net_impl.cuh"
THB_DEFINE_SHAPE(3, 4, 3)
