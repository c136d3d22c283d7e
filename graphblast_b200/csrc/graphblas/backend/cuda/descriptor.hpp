// graphblast_b200 backend — Descriptor: per-call flags, CLI knobs and the
// grow-only device scratch every operation borrows from.
//
// Replaces reference graphblas/backend/cuda/descriptor.hpp:14-287.  Same field
// table (desc_[GrB_NDESCFIELD]), same toggle() rule (:141-154), same knob names
// and accessors, same loadArgs() mapping from po::variables_map (:207-287).
// Data members the reference drivers reach through `#define private public`
// keep their names: max_niter_, timing_, lastmxv_, debug_
// (reference algorithm/bfs.hpp:46,54,56).
//
// B200-first differences:
//  * no moderngpu context; scratch is a set of typed arenas sized in size_t
//    (the reference sizes scratch in `int` and overflows at RMAT-24 unless
//    --memusage <= 0.5, reference spmspv.hpp:60-66);
//  * push-direction accumulator + touched-bitmap arenas keep an invariant
//    ("all identity" / "all zero") between calls so no O(n) clear per level.
#ifndef GRAPHBLAS_BACKEND_CUDA_DESCRIPTOR_HPP_
#define GRAPHBLAS_BACKEND_CUDA_DESCRIPTOR_HPP_

#include <vector>
#include <string>

#include "graphblas/backend/cuda/util.hpp"

namespace graphblas {
namespace backend {

// Scratch arenas (each grows independently, never shrinks).
enum ScratchSlot {
  GB_SCRATCH_ACC = 0,     // push: dense accumulator, one value per output vertex
  GB_SCRATCH_BITS,        // push: touched bitmap, one bit per output vertex
  GB_SCRATCH_OFFS,        // push: scanned frontier degrees
  GB_SCRATCH_BLOCKSUM,    // compaction: per-CTA counts / offsets
  GB_SCRATCH_COUNTERS,    // small device counters (64 x 8 bytes)
  GB_SCRATCH_CARRY_ROW,   // pull: per-CTA carry-out row ids
  GB_SCRATCH_CARRY_VAL,   // pull: per-CTA carry-out partials
  GB_SCRATCH_VEC_A,       // generic n-sized temporaries
  GB_SCRATCH_VEC_B,
  GB_SCRATCH_CUB,         // cub temp storage
  GB_SCRATCH_LOOKBACK,    // compaction: ticket cell + per-CTA look-back status
  GB_SCRATCH_BFS,         // fused BFS: visited x2, frontier, next bitmaps + cells
  GB_SCRATCH_NSLOTS
};

class Descriptor {
 public:
  Descriptor() : desc_{ GrB_DEFAULT, GrB_DEFAULT, GrB_DEFAULT, GrB_DEFAULT,
    GrB_FIXEDROW, GrB_32, GrB_32, GrB_128, GrB_PUSHPULL, GrB_16, GrB_CUDA},
    d_buffer_(NULL), d_buffer_size_(0), d_temp_(NULL), d_temp_size_(0),
    ta_(0), tb_(0), mode_(""), split_(0),
    enable_split_(0), niter_(0), max_niter_(0), directed_(0), timing_(0),
    transpose_(0), mtxinfo_(0), verbose_(0), mxvmode_(0),
    lastmxv_(GrB_PUSHONLY), switchpoint_(0), dirinfo_(0), struconly_(0),
    opreuse_(0), memusage_(0), endbit_(0), sort_(0), atomic_(0),
    earlyexit_(0), fusedmask_(0), nthread_(0), ndevice_(0), debug_(0),
    memory_(0), acc_elems_(0), acc_identity_bits_(0), acc_elem_bytes_(0),
    acc_valid_(false), bits_words_(0), bits_valid_(false),
    lookback_epoch_(0), lookback_ticket_(0) {
    for (int i = 0; i < GB_SCRATCH_NSLOTS; ++i) {
      slot_ptr_[i]  = NULL;
      slot_size_[i] = 0;
    }
  }

  ~Descriptor();

  // C API Methods
  Info set(Desc_field field, Desc_value  value);
  Info get(Desc_field field, Desc_value* value) const;

  // Useful methods
  Info toggle(Desc_field field);
  Info loadArgs(const po::variables_map& vm);

  inline bool debug()  { return debug_;  }
  inline bool memory() { return memory_; }

  inline bool struconly()    { return struconly_; }
  inline bool split()        { return split_ && enable_split_; }
  inline bool dirinfo()      { return dirinfo_; }
  inline bool earlyexit()    { return earlyexit_; }
  inline bool opreuse()      { return opreuse_; }
  inline bool endbit()       { return endbit_; }
  inline bool sort()         { return sort_; }
  inline bool fusedmask()    { return fusedmask_; }
  inline bool atomic()       { return atomic_; }
  inline float switchpoint() { return switchpoint_; }
  inline float memusage()    { return memusage_; }

 public:  // (private in the reference; its drivers `#define private public`)
  // Legacy two-buffer interface (reference descriptor.hpp:156-192), kept for
  // callers that size "buffer"/"temp" by name.
  Info resize(size_t target, std::string field);
  Info clear(std::string field);

 public:
  // Arena interface used by this backend's operations.
  void* scratch(ScratchSlot slot, size_t bytes) {
    if (bytes > slot_size_[slot]) {
      if (slot_ptr_[slot] != NULL) {
        // In-flight kernels may still read the old block.
        CUDA_CALL(cudaStreamSynchronize(gbStream()));
        CUDA_CALL(cudaFree(slot_ptr_[slot]));
      }
      size_t want = bytes + bytes/4 + 256;
      CUDA_CALL(cudaMalloc(&slot_ptr_[slot], want));
      slot_size_[slot] = want;
      if (slot == GB_SCRATCH_ACC)  acc_valid_  = false;
      if (slot == GB_SCRATCH_BITS) bits_valid_ = false;
    }
    return slot_ptr_[slot];
  }
  size_t scratchSize(ScratchSlot slot) const { return slot_size_[slot]; }

  // Look-back state of the single-pass compaction: cell 0 is a ticket counter
  // that only ever grows, cells 1..nblocks hold (epoch, flag, value) words.  The
  // block is zeroed when (re)allocated; afterwards nothing is ever reset — each
  // launch uses a fresh epoch and the ticket base the host has kept count of.
  unsigned long long* lookback(size_t nblocks) {
    const size_t bytes = (nblocks + 1)*sizeof(unsigned long long);
    if (bytes > slot_size_[GB_SCRATCH_LOOKBACK]) {
      void* p = scratch(GB_SCRATCH_LOOKBACK, bytes);
      CUDA_CALL(cudaMemsetAsync(p, 0, slot_size_[GB_SCRATCH_LOOKBACK], gbStream()));
      lookback_epoch_  = 0;
      lookback_ticket_ = 0;
    }
    return reinterpret_cast<unsigned long long*>(
        slot_ptr_[GB_SCRATCH_LOOKBACK]);
  }
  unsigned int       lookback_epoch_;
  unsigned long long lookback_ticket_;

  // Device counters: 64 x 8-byte cells, zero when first handed out.  Cell 2 is
  // the "finished CTAs" counter of the compaction's count pass, which leaves it
  // at zero again.
  unsigned long long* counters() {
    if (slot_ptr_[GB_SCRATCH_COUNTERS] == NULL) {
      void* p = scratch(GB_SCRATCH_COUNTERS, 64*sizeof(unsigned long long));
      CUDA_CALL(cudaMemsetAsync(p, 0, slot_size_[GB_SCRATCH_COUNTERS], gbStream()));
    }
    return reinterpret_cast<unsigned long long*>(
        slot_ptr_[GB_SCRATCH_COUNTERS]);
  }

 public:  // (private in the reference; its drivers `#define private public`)
  Desc_value desc_[GrB_NDESCFIELD];

  void*       d_buffer_;      // legacy scratch
  size_t      d_buffer_size_;
  void*       d_temp_;        // legacy cub scratch
  size_t      d_temp_size_;

  void*       slot_ptr_[GB_SCRATCH_NSLOTS];
  size_t      slot_size_[GB_SCRATCH_NSLOTS];

  // Algorithm specific params
  int         ta_;
  int         tb_;
  std::string mode_;
  bool        split_;
  bool        enable_split_;

  // General params
  int         niter_;
  int         max_niter_;
  int         directed_;
  int         timing_;
  bool        transpose_;
  bool        mtxinfo_;
  bool        verbose_;

  // mxv params
  int         mxvmode_;
  Desc_value  lastmxv_;
  float       switchpoint_;
  bool        dirinfo_;
  bool        struconly_;
  bool        opreuse_;

  // mxv (spmspv/push) params
  float       memusage_;
  bool        endbit_;
  bool        sort_;
  bool        atomic_;

  // mxv (spmv/pull) params
  bool        earlyexit_;
  bool        fusedmask_;

  // GPU params
  int         nthread_;
  int         ndevice_;
  bool        debug_;
  bool        memory_;

  // State of the push accumulator arena: which identity it is filled with.
  size_t      acc_elems_;
  unsigned    acc_identity_bits_;
  size_t      acc_elem_bytes_;
  bool        acc_valid_;
  size_t      bits_words_;
  bool        bits_valid_;
};

inline Descriptor::~Descriptor() {
  if (d_buffer_ != NULL) cudaFree(d_buffer_);
  if (d_temp_   != NULL) cudaFree(d_temp_);
  for (int i = 0; i < GB_SCRATCH_NSLOTS; ++i)
    if (slot_ptr_[i] != NULL) cudaFree(slot_ptr_[i]);
}

inline Info Descriptor::set(Desc_field field, Desc_value value) {
  desc_[field] = value;
  return GrB_SUCCESS;
}

inline Info Descriptor::get(Desc_field field, Desc_value* value) const {
  *value = desc_[field];
  return GrB_SUCCESS;
}

// Fields 0..3 (MASK, OUTP, INP0, INP1) flip between GrB_DEFAULT and their one
// non-default value; the enum is laid out so that value == field for the first
// three (GrB_SCMP=0, GrB_REPLACE=1, GrB_TRAN=2) and INP1 also maps to GrB_TRAN.
inline Info Descriptor::toggle(Desc_field field) {
  int idx = static_cast<int>(field);
  if (idx >= 4) return GrB_SUCCESS;
  if (desc_[field] != GrB_DEFAULT)
    desc_[field] = GrB_DEFAULT;
  else
    desc_[field] = (idx == 3) ? GrB_TRAN : static_cast<Desc_value>(idx);
  return GrB_SUCCESS;
}

inline Info Descriptor::resize(size_t target, std::string field) {
  void**  ptr  = (field == "temp") ? &d_temp_      : &d_buffer_;
  size_t* size = (field == "temp") ? &d_temp_size_ : &d_buffer_size_;
  if (target > *size) {
    void* fresh = NULL;
    CUDA_CALL(cudaMalloc(&fresh, target));
    if (*ptr != NULL) {
      CUDA_CALL(cudaMemcpyAsync(fresh, *ptr, *size, cudaMemcpyDeviceToDevice, gbStream()));
      CUDA_CALL(cudaStreamSynchronize(gbStream()));
      CUDA_CALL(cudaFree(*ptr));
    }
    *ptr  = fresh;
    *size = target;
  }
  return GrB_SUCCESS;
}

inline Info Descriptor::clear(std::string field) {
  if (field == "buffer" && d_buffer_ != NULL)
    CUDA_CALL(cudaMemsetAsync(d_buffer_, 0, d_buffer_size_, gbStream()));
  else if (field == "temp" && d_temp_ != NULL)
    CUDA_CALL(cudaMemsetAsync(d_temp_, 0, d_temp_size_, gbStream()));
  return GrB_SUCCESS;
}

inline Info Descriptor::loadArgs(const po::variables_map& vm) {
  // Algorithm specific params
  ta_             = vm["ta"            ].as<int>();
  tb_             = vm["tb"            ].as<int>();
  mode_           = vm["mode"          ].as<std::string>();
  split_          = vm["split"         ].as<bool>();

  // General params
  niter_          = vm["niter"         ].as<int>();
  max_niter_      = vm["max_niter"     ].as<int>();
  directed_       = vm["directed"      ].as<int>();
  timing_         = vm["timing"        ].as<int>();
  transpose_      = vm["transpose"     ].as<bool>();
  mtxinfo_        = vm["mtxinfo"       ].as<bool>();
  verbose_        = vm["verbose"       ].as<bool>();

  // mxv params
  mxvmode_        = vm["mxvmode"       ].as<int>();
  switchpoint_    = vm["switchpoint"   ].as<float>();
  dirinfo_        = vm["dirinfo"       ].as<bool>();
  struconly_      = vm["struconly"     ].as<bool>();
  opreuse_        = vm["opreuse"       ].as<bool>();

  // mxv (spmspv/push) params
  memusage_       = vm["memusage"      ].as<float>();
  endbit_         = vm["endbit"        ].as<bool>();
  sort_           = vm["sort"          ].as<bool>();
  atomic_         = vm["atomic"        ].as<bool>();

  // mxv (spmv/pull) params
  earlyexit_      = vm["earlyexit"     ].as<bool>();
  fusedmask_      = vm["fusedmask"     ].as<bool>();

  // GPU params
  nthread_        = vm["nthread"       ].as<int>();
  ndevice_        = vm["ndevice"       ].as<int>();
  debug_          = vm["debug"         ].as<bool>();
  memory_         = vm["memory"        ].as<bool>();

  switch (mxvmode_) {
    case 0: CHECK(set(GrB_MXVMODE, GrB_PUSHPULL)); break;
    case 1: CHECK(set(GrB_MXVMODE, GrB_PUSHONLY)); break;
    case 2: CHECK(set(GrB_MXVMODE, GrB_PULLONLY)); break;
    default: std::cout << "Error: incorrect mxvmode selection!\n";
  }

  switch (nthread_) {
    case 32:   CHECK(set(GrB_NT, GrB_32));   break;
    case 64:   CHECK(set(GrB_NT, GrB_64));   break;
    case 128:  CHECK(set(GrB_NT, GrB_128));  break;
    case 256:  CHECK(set(GrB_NT, GrB_256));  break;
    case 512:  CHECK(set(GrB_NT, GrB_512));  break;
    case 1024: CHECK(set(GrB_NT, GrB_1024)); break;
    default: std::cout << "Error: incorrect nthread selection!\n";
  }

  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_DESCRIPTOR_HPP_
