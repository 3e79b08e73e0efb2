// Tile-kernel instantiations for tensor-product shape (3,4,1) = degrees (2,3,0); generated list, see thb_tile.cu
#include "thb_net_impl.cuh"
THB_DEFINE_SHAPE(3, 4, 1)
