// Tile-kernel instantiations for tensor-product shape (2,2,1) = degrees (1,1,0); generated list, see thb_tile.cu
#include "thb_net_impl.cuh"
THB_DEFINE_SHAPE(2, 2, 1)
