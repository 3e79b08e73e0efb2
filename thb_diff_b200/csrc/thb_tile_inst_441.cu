// Tile-kernel instantiations for tensor-product shape (4,4,1) = degrees (3,3,0); generated list, see thb_tile.cu
#include "thb_net_impl.cuh"
THB_DEFINE_SHAPE(4, 4, 1)
