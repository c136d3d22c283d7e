// graphblast_b200 backend — per-matrix index of the hub-cached pull SpMV
// (kernels/spmv_hub.cuh, spmv_hub.hpp).  Kept in its own header because the
// matrix container caches one per traversed direction.
#ifndef GRAPHBLAS_BACKEND_CUDA_HUB_INDEX_HPP_
#define GRAPHBLAS_BACKEND_CUDA_HUB_INDEX_HPP_

#include "graphblas/backend/cuda/util.hpp"

namespace graphblas {
namespace backend {

// Per (matrix, direction): which columns are hubs, the encoded column array, the
// compact list of non-empty rows and the tile records.
struct HubIndex {
  Index*         enc_ci;      // [nnz + 8] column id, or GB_HUB_FLAG | slot
  Index*         hub_ids;     // [capacity] column id of every slot (first `count`)
  void*          hub_vals;    // [capacity] x 4 bytes, refreshed per call
  Index*         ne_ptr;      // [m + PAD] first nonzero of every non-empty row
  Index*         ne_rows;     // [m + PAD] its row id
  Index*         empty_rows;  // [nempty]
  int4*          desc;        // [ntiles] tile records
  unsigned char* chunk_rel;   // local row of the first nonzero of every 8-entry chunk
  Index          m;           // non-empty rows
  Index          nempty;
  int            ntiles;
  int            count;       // hub slots in use
  int            capacity;
  double         coverage;    // share of the stored entries that reference a hub
  const Index*   key;         // colind pointer this was built from
  Index          key_nvals;
  HubIndex() : enc_ci(NULL), hub_ids(NULL), hub_vals(NULL), ne_ptr(NULL),
               ne_rows(NULL), empty_rows(NULL), desc(NULL), chunk_rel(NULL), m(0),
               nempty(0), ntiles(0), count(0), capacity(0), coverage(0.),
               key(NULL), key_nvals(-1) {}
  void release() {
    void* all[] = {enc_ci, hub_ids, hub_vals, ne_ptr, ne_rows, empty_rows, desc,
                   chunk_rel};
    for (size_t i = 0; i < sizeof(all)/sizeof(all[0]); ++i)
      if (all[i] != NULL) gbFree(all[i]);
    *this = HubIndex();
  }
};

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_HUB_INDEX_HPP_
