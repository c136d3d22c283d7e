// graphblast_b200 backend — operation dispatch: storage-type x direction decision
// tree behind every frontend template in graphblas/operations.hpp.
//
// Replaces reference graphblas/backend/cuda/operations.hpp:18-1435.  Template
// parameter orders match the explicit instantiations the frontend writes
// (backend::mxm<c,a,b,m>, vxm<W,U,a,M>, mxv<W,U,a,M>, applyVxm<W,U,a,M>;
// reference graphblas/operations.hpp:47,85,125,863).  Operations that no
// hot-path algorithm reaches are declared and return GrB_NOT_IMPLEMENTED
// ("must declare").
//
// Direction choice for vxm/mxv (reference :124-139, :252-266):
//   CSR-only non-symmetric matrix -> vxm forced to push, mxv forced to pull;
//   GrB_PUSHPULL  -> Vector::convert() heuristic on the input vector;
//   GrB_PUSHONLY / GrB_PULLONLY -> input converted if needed.
// vxm is executed as mxv on the transposed matrix by toggling GrB_INP1 for the
// duration of the call; desc->lastmxv_ records the direction taken.
#ifndef GRAPHBLAS_BACKEND_CUDA_OPERATIONS_HPP_
#define GRAPHBLAS_BACKEND_CUDA_OPERATIONS_HPP_

#include <vector>
#include <typeinfo>

#include "graphblas/backend/cuda/vector.hpp"
#include "graphblas/backend/cuda/matrix.hpp"
#include "graphblas/backend/cuda/spmv.hpp"
#include "graphblas/backend/cuda/spmspv.hpp"
#include "graphblas/backend/cuda/spgemm.hpp"
#include "graphblas/backend/cuda/ewiseadd.hpp"
#include "graphblas/backend/cuda/ewisemult.hpp"
#include "graphblas/backend/cuda/assign.hpp"
#include "graphblas/backend/cuda/reduce.hpp"
#include "graphblas/backend/cuda/apply.hpp"
#include "graphblas/backend/cuda/tri.hpp"

namespace graphblas {
namespace backend {

template <typename c, typename a, typename b, typename m,
          typename BinaryOpT,     typename SemiringT>
Info mxm(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Matrix<b>* B, Descriptor* desc) {
  Storage A_mat_type;
  Storage B_mat_type;
  CHECK(A->getStorage(&A_mat_type));
  CHECK(B->getStorage(&B_mat_type));

  if (A_mat_type == GrB_SPARSE && B_mat_type == GrB_SPARSE) {
    CHECK(C->setStorage(GrB_SPARSE));
    if (mask) {
      CHECK(spgemmMasked(&C->sparse_, mask, accum, op, &A->sparse_, &B->sparse_, desc));
    } else {
      std::cout << "Error: Unmasked SpGEMM not implemented yet!\n";
      return GrB_NOT_IMPLEMENTED;
    }
  } else {
    std::cout << "Error: SpMM and GEMM not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  return GrB_SUCCESS;
}

// Shared body of vxm / mxv once the descriptor says which side is transposed.
template <bool IsVxm, typename W, typename U, typename a, typename M,
          typename BinaryOpT, typename SemiringT>
Info mxvDispatch(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Vector<U>* u, Descriptor* desc) {
  Vector<U>* u_t = const_cast<Vector<U>*>(u);

  Storage u_vec_type;
  Storage A_mat_type;
  CHECK(u->getStorage(&u_vec_type));
  CHECK(A->getStorage(&A_mat_type));

  LoadBalanceMode lb_mode = getEnv("GRB_LOAD_BALANCE_MODE",
      GrB_LOAD_BALANCE_MERGE);

  SparseMatrixFormat A_format;
  bool A_symmetric;
  CHECK(A->getFormat(&A_format));
  CHECK(A->getSymmetry(&A_symmetric));

  Desc_value mxv_mode;
  CHECK(desc->get(GrB_MXVMODE, &mxv_mode));

  // Conversions of the input vector decide the direction.
  if (!A_symmetric && A_format == GrB_SPARSE_MATRIX_CSRONLY) {
    if (IsVxm) {
      if (u_vec_type == GrB_DENSE)
        CHECK(u_t->dense2sparse(op.identity(), desc));
    } else {
      if (u_vec_type == GrB_SPARSE)
        CHECK(u_t->sparse2dense(op.identity(), desc));
    }
  } else if (mxv_mode == GrB_PUSHPULL) {
    CHECK(u_t->convert(op.identity(), desc->switchpoint(), desc));
  } else if (mxv_mode == GrB_PUSHONLY && u_vec_type == GrB_DENSE) {
    CHECK(u_t->dense2sparse(op.identity(), desc));
  } else if (mxv_mode == GrB_PULLONLY && u_vec_type == GrB_SPARSE) {
    CHECK(u_t->sparse2dense(op.identity(), desc));
  }
  CHECK(u->getStorage(&u_vec_type));

  bool run_pull = !(A_mat_type == GrB_SPARSE && u_vec_type == GrB_SPARSE);
  if (!run_pull) {
    if (lb_mode == GrB_LOAD_BALANCE_MERGE) {
      // w becomes a sparse vector only if the push really runs: on a hand-back
      // its previous storage (and with it a dense w that accum combines into) must
      // survive untouched, so the tag is restored below.
      Storage w_before;
      CHECK(w->getStorage(&w_before));
      CHECK(w->setStorage(GrB_SPARSE));
      // In the automatic mode the push may hand the call back when the frontier
      // owns too many of the edges (spmspv.hpp); both orientations must exist.
      bool prefer_pull = false;
      const bool may_switch = (mxv_mode == GrB_PUSHPULL) &&
          (A_symmetric || A_format == GrB_SPARSE_MATRIX_CSRCSC);
      CHECK(spmspvMerge(&w->sparse_, mask, accum, op,
          &A->sparse_, &u->sparse_, desc, may_switch ? &prefer_pull : NULL));
      if (prefer_pull) {
        w->vec_type_ = w_before;
        CHECK(u_t->sparse2dense(op.identity(), desc));
        run_pull = true;
      }
    } else if (lb_mode == GrB_LOAD_BALANCE_SIMPLE) {
      std::cout << "Simple SPMSPV not implemented yet!\n";
      return GrB_NOT_IMPLEMENTED;
    } else if (lb_mode == GrB_LOAD_BALANCE_TWC) {
      std::cout << "Error: B40C load-balance algorithm not implemented yet!\n";
      return GrB_NOT_IMPLEMENTED;
    } else {
      std::cout << "Error: Invalid load-balance algorithm!\n";
    }
    if (!run_pull) desc->lastmxv_ = GrB_PUSHONLY;
  }
  if (run_pull) {
    if (IsVxm) CHECK(w->setStorage(GrB_DENSE));
    else       CHECK(w->sparse2dense(op.identity(), desc));
    if (A_mat_type == GrB_SPARSE) {
      CHECK(spmv(&w->dense_, mask, accum, op, &A->sparse_, &u->dense_, desc));
    } else {
      std::cout << "Error: GEMV not implemented yet!\n";
      return GrB_NOT_IMPLEMENTED;
    }
    desc->lastmxv_ = GrB_PULLONLY;
  }
  return GrB_SUCCESS;
}

template <typename W, typename U, typename a, typename M,
          typename BinaryOpT, typename SemiringT>
Info vxm(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Matrix<a>* A, Descriptor* desc) {
  if (desc->debug()) {
    std::cout << "===Begin vxm===\n";
    CHECK(const_cast<Vector<U>*>(u)->print());
  }

  Desc_value inp0_mode;
  CHECK(desc->get(GrB_INP0, &inp0_mode));
  if (inp0_mode != GrB_DEFAULT) return GrB_INVALID_VALUE;

  // Treat vxm as an mxv with transposed matrix
  CHECK(desc->toggle(GrB_INP1));
  Info err = mxvDispatch<true>(w, mask, accum, op, A, u, desc);
  CHECK(desc->toggle(GrB_INP1));
  if (err != GrB_SUCCESS) return err;

  if (desc->debug()) {
    std::cout << "===End vxm===\n";
    CHECK(w->print());
  }
  return GrB_SUCCESS;
}

template <typename W, typename a, typename U, typename M,
          typename BinaryOpT, typename SemiringT>
Info mxv(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Vector<U>* u, Descriptor* desc) {
  if (desc->debug()) {
    std::cout << "===Begin mxv===\n";
    CHECK(const_cast<Vector<U>*>(u)->print());
  }

  Desc_value inp1_mode;
  CHECK(desc->get(GrB_INP1, &inp1_mode));
  if (inp1_mode != GrB_DEFAULT) return GrB_INVALID_VALUE;

  CHECK((mxvDispatch<false>(w, mask, accum, op, A, u, desc)));

  if (desc->debug()) {
    std::cout << "===End mxv===\n";
    CHECK(w->print());
  }
  return GrB_SUCCESS;
}

template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Vector<V>* v, Descriptor* desc) {
  Vector<V>* v_t = const_cast<Vector<V>*>(v);
  CHECK(u->materialize());
  CHECK(v->materialize());
  CHECK(w->materialize());
  if (mask != NULL) CHECK(mask->materialize());

  Storage u_vec_type;
  Storage v_vec_type;
  CHECK(u->getStorage(&u_vec_type));
  CHECK(v->getStorage(&v_vec_type));

  // sparse x sparse: the reference flips v's tag to dense (operations.hpp:365-371)
  if (u_vec_type == GrB_SPARSE && v_vec_type == GrB_SPARSE)
    CHECK(v_t->setStorage(GrB_DENSE));
  CHECK(u->getStorage(&u_vec_type));
  CHECK(v->getStorage(&v_vec_type));

  if (u_vec_type == GrB_DENSE && v_vec_type == GrB_DENSE) {
    if (mask != NULL) {
      Storage mask_type;
      CHECK(mask->getStorage(&mask_type));
      if (mask_type == GrB_DENSE) {
        CHECK(w->setStorage(GrB_DENSE));
        CHECK(eWiseMultInner(&w->dense_, mask, accum, op,
            &u->dense_, &v->dense_, desc));
      } else if (mask_type == GrB_SPARSE) {
        CHECK(w->setStorage(GrB_SPARSE));
        CHECK(eWiseMultInner(&w->sparse_,
            &mask->sparse_, accum, op, &u->dense_, &v->dense_, desc));
      } else {
        return GrB_INVALID_OBJECT;
      }
    } else {
      CHECK(w->setStorage(GrB_DENSE));
      CHECK(eWiseMultInner(&w->dense_, mask, accum, op, &u->dense_, &v->dense_, desc));
    }
  } else if (u_vec_type == GrB_SPARSE && v_vec_type == GrB_DENSE) {
    CHECK(w->setStorage(GrB_SPARSE));
    CHECK(eWiseMultInner(&w->sparse_, mask, accum, op,
        &u->sparse_, &v->dense_, false, desc));
  } else if (u_vec_type == GrB_DENSE && v_vec_type == GrB_SPARSE) {
    CHECK(w->setStorage(GrB_SPARSE));
    CHECK(eWiseMultInner(&w->sparse_, mask, accum, op,
        &v->sparse_, &u->dense_, true, desc));
  } else {
    return GrB_INVALID_OBJECT;
  }
  return GrB_SUCCESS;
}

template <typename c, typename a, typename b, typename m,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Matrix<b>* B, Descriptor* desc) {
  std::cout << "Error: eWiseMult matrix variant not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

// Extension: matrix (x) broadcast scalar
template <typename c, typename a, typename b, typename m,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, b val, Descriptor* desc) {
  Storage A_mat_type;
  CHECK(A->getStorage(&A_mat_type));
  if (A_mat_type != GrB_SPARSE) {
    std::cout << "eWiseMult Dense Matrix Broadcast Scalar\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return (A_mat_type == GrB_DENSE) ? GrB_NOT_IMPLEMENTED : GrB_INVALID_OBJECT;
  }
  if (mask != NULL) {
    std::cout << "eWiseMult Sparse Matrix Broadcast Scalar with Mask\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  CHECK(C->setStorage(GrB_SPARSE));
  CHECK(eWiseMultInner(&C->sparse_, mask, accum, op, &A->sparse_, val, desc));
  return GrB_SUCCESS;
}

// Extension: matrix (x) broadcast vector (column vector; row vector when
// GrB_INP1 is GrB_TRAN)
template <typename c, typename a, typename b, typename m,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Vector<b>* B, Descriptor* desc) {
  Desc_value inp0_mode, inp1_mode;
  CHECK(desc->get(GrB_INP0, &inp0_mode));
  CHECK(desc->get(GrB_INP1, &inp1_mode));
  if (inp0_mode != GrB_DEFAULT) return GrB_INVALID_VALUE;
  CHECK(B->materialize());

  Storage A_mat_type;
  Storage B_vec_type;
  CHECK(A->getStorage(&A_mat_type));
  CHECK(B->getStorage(&B_vec_type));

  if (A_mat_type != GrB_SPARSE) {
    std::cout << "eWiseMult Dense Matrix Broadcast Vector\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return (A_mat_type == GrB_DENSE) ? GrB_NOT_IMPLEMENTED : GrB_INVALID_OBJECT;
  }
  if (mask != NULL) {
    std::cout << "eWiseMult Sparse Matrix Broadcast Vector with Mask\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  CHECK(C->setStorage(GrB_SPARSE));
  if (B_vec_type == GrB_SPARSE) {
    std::cout << "eWiseMult Sparse Matrix Broadcast Sparse Vector\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  if (inp1_mode != GrB_TRAN)
    CHECK(eWiseMultColInner(&C->sparse_, mask, accum, op,
        &A->sparse_, &B->dense_, desc));
  else
    CHECK(eWiseMultRowInner(&C->sparse_, mask, accum, op,
        &A->sparse_, &B->dense_, desc));
  return GrB_SUCCESS;
}

template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAdd(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Vector<V>* v, Descriptor* desc) {
  Vector<U>* u_t = const_cast<Vector<U>*>(u);
  Vector<V>* v_t = const_cast<Vector<V>*>(v);
  CHECK(u->materialize());
  CHECK(v->materialize());
  CHECK(w->materialize());
  if (mask != NULL) CHECK(mask->materialize());

  Storage u_vec_type;
  Storage v_vec_type;
  CHECK(u->getStorage(&u_vec_type));
  CHECK(v->getStorage(&v_vec_type));

  // An in-place sparse operand is densified first (reference :598-607).
  const void* w_addr = reinterpret_cast<const void*>(w);
  if (reinterpret_cast<const void*>(u) == w_addr && u_vec_type == GrB_SPARSE) {
    u_t->sparse2dense(op.identity(), desc);
    u_vec_type = GrB_DENSE;
  } else if (reinterpret_cast<const void*>(v) == w_addr &&
             v_vec_type == GrB_SPARSE) {
    v_t->sparse2dense(op.identity(), desc);
    v_vec_type = GrB_DENSE;
  }

  CHECK(w->setStorage(GrB_DENSE));
  if (u_vec_type == GrB_SPARSE && v_vec_type == GrB_SPARSE) {
    CHECK(eWiseAddInner(&w->dense_, mask, accum, op, &u->sparse_, &v->sparse_, desc));
  } else if (u_vec_type == GrB_DENSE && v_vec_type == GrB_DENSE) {
    CHECK(eWiseAddInner(&w->dense_, mask, accum, op, &u->dense_, &v->dense_, desc));
  } else if (u_vec_type == GrB_SPARSE && v_vec_type == GrB_DENSE) {
    CHECK(eWiseAddInner(&w->dense_, mask, accum, op,
        &u->sparse_, &v->dense_, false, desc));
  } else if (u_vec_type == GrB_DENSE && v_vec_type == GrB_SPARSE) {
    CHECK(eWiseAddInner(&w->dense_, mask, accum, op,
        &v->sparse_, &u->dense_, true, desc));
  } else {
    std::cout << "Error: eWiseAdd backend invalid choice!\n";
    return GrB_INVALID_OBJECT;
  }
  return GrB_SUCCESS;
}

template <typename c, typename a, typename b, typename m,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAdd(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Matrix<b>* B, Descriptor* desc) {
  std::cout << "Error: eWiseAdd matrix variant not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

// Extension: vector (+) broadcast scalar
template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAdd(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, V val, Descriptor* desc) {
  CHECK(u->materialize());
  CHECK(w->materialize());
  Storage u_vec_type;
  CHECK(u->getStorage(&u_vec_type));
  if (u_vec_type != GrB_DENSE && u_vec_type != GrB_SPARSE)
    return GrB_INVALID_OBJECT;
  if (mask != NULL) {
    std::cout << "eWiseAdd Vector-Scalar with Mask\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  CHECK(w->setStorage(GrB_DENSE));
  if (u_vec_type == GrB_DENSE)
    CHECK(eWiseAddInner(&w->dense_, mask, accum, op, &u->dense_, val, desc));
  else
    CHECK(eWiseAddInner(&w->dense_, mask, accum, op, &u->sparse_, val, desc));
  return GrB_SUCCESS;
}

template <typename W, typename U, typename M,
          typename BinaryOpT>
Info extract(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, const Vector<U>* u,
    const std::vector<Index>* indices, Index nindices, Descriptor* desc) {
  std::cout << "Error: extract vector variant not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename W, typename U, typename M,
          typename BinaryOpT>
Info assignIndexed(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    const Vector<U>* u, int* indices, Index nindices, Descriptor* desc) {
  std::cout << "Error: assignIndexed not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

// Masked constant assign
template <typename W, typename T, typename M,
          typename BinaryOpT>
Info assign(Vector<W>* w, Vector<M>* mask, BinaryOpT accum, T val,
    const Vector<Index>* indices, Index nindices, Descriptor* desc) {
  if (desc->debug()) {
    std::cout << "===Begin assign===\n";
    std::cout << "Input: " << val << std::endl;
  }

  Storage vec_type;
  CHECK(w->getStorage(&vec_type));
  // The target is written in part; a dense mask is read through its bitmap
  // shadow when that is current (which lazily held values imply), except by the
  // sparse-target filter, which reads mask values.
  CHECK(w->materialize());
  if (vec_type == GrB_SPARSE && mask != NULL) CHECK(mask->materialize());

  if (vec_type == GrB_SPARSE) {
    CHECK(assignSparse(&w->sparse_, mask, accum, val, indices, nindices, desc));
  } else if (vec_type == GrB_DENSE) {
    CHECK(assignDense(&w->dense_, mask, accum, val, indices, nindices, desc));
  }

  if (desc->debug()) {
    std::cout << "===End assign===\n";
    CHECK(w->print());
  }
  return GrB_SUCCESS;
}

template <typename W, typename U, typename M,
          typename BinaryOpT,     typename UnaryOpT>
Info apply(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, UnaryOpT op,
    const Vector<U>* u, Descriptor* desc) {
  Vector<U>* u_t = const_cast<Vector<U>*>(u);
  CHECK(u->materialize());
  CHECK(w->materialize());
  if (mask != NULL) CHECK(mask->materialize());
  Storage u_vec_type;
  CHECK(u->getStorage(&u_vec_type));
  if (u_vec_type == GrB_SPARSE) {
    CHECK(w->setStorage(GrB_SPARSE));
    applySparse(&w->sparse_, mask, accum, op, &u_t->sparse_, desc);
  } else if (u_vec_type == GrB_DENSE) {
    CHECK(w->setStorage(GrB_DENSE));
    applyDense(&w->dense_, mask, accum, op, &u_t->dense_, desc);
  } else {
    return GrB_UNINITIALIZED_OBJECT;
  }
  return GrB_SUCCESS;
}

template <typename c, typename a, typename m,
          typename BinaryOpT,     typename UnaryOpT>
Info apply(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, UnaryOpT op,
    const Matrix<a>* A, Descriptor* desc) {
  Matrix<a>* A_t = const_cast<Matrix<a>*>(A);
  Storage A_mat_type;
  CHECK(A->getStorage(&A_mat_type));
  if (A_mat_type == GrB_SPARSE) {
    CHECK(C->setStorage(GrB_SPARSE));
    applySparse(&C->sparse_, mask, accum, op, &A_t->sparse_, desc);
  } else if (A_mat_type == GrB_DENSE) {
    CHECK(C->setStorage(GrB_DENSE));
    applyDense(&C->dense_, mask, accum, op, &A_t->dense_, desc);
  } else {
    return GrB_UNINITIALIZED_OBJECT;
  }
  return GrB_SUCCESS;
}

// matrix rows -> vector
template <typename W, typename a, typename M,
          typename BinaryOpT,     typename MonoidT>
Info reduce(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, MonoidT op,
    const Matrix<a>* A, Descriptor* desc) {
  Storage mat_type;
  CHECK(A->getStorage(&mat_type));
  CHECK(w->setStorage(GrB_DENSE));

  if (mask != NULL) {
    std::cout << "Error: Masked reduce not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  if (mat_type == GrB_SPARSE)
    CHECK(reduceInner(&w->dense_, mask, accum, op, &A->sparse_, desc));
  else if (mat_type == GrB_DENSE)
    CHECK(reduceInner(&w->dense_, mask, accum, op, &A->dense_, desc));
  else
    return GrB_UNINITIALIZED_OBJECT;
  return GrB_SUCCESS;
}

// vector -> scalar
template <typename T, typename U,
          typename BinaryOpT, typename MonoidT>
Info reduce(T* val, BinaryOpT accum, MonoidT op, const Vector<U>* u, Descriptor* desc) {
  Storage vec_type;
  CHECK(u->getStorage(&vec_type));

  if (vec_type == GrB_SPARSE)
    CHECK(reduceInner(val, accum, op, &u->sparse_, desc));
  else if (vec_type == GrB_DENSE)
    CHECK(reduceInner(val, accum, op, &u->dense_, desc));
  else
    return GrB_UNINITIALIZED_OBJECT;

  if (desc->debug())
    std::cout << "reduce output: " << *val << std::endl;
  return GrB_SUCCESS;
}

// matrix -> scalar
template <typename T, typename a,
          typename BinaryOpT,     typename MonoidT>
Info reduce(T* val, BinaryOpT accum, MonoidT op, const Matrix<a>* A, Descriptor* desc) {
  Storage mat_type;
  CHECK(A->getStorage(&mat_type));

  if (mat_type == GrB_SPARSE) {
    CHECK(reduceInner(val, accum, op, &A->sparse_, desc));
  } else if (mat_type == GrB_DENSE) {
    std::cout << "Error: reduce matrix-scalar for dense matrix\n";
    std::cout << "not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  } else {
    return GrB_UNINITIALIZED_OBJECT;
  }
  return GrB_SUCCESS;
}

template <typename c, typename a, typename m,
          typename BinaryOpT>
Info transpose(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, const Matrix<a>* A,
    Descriptor* desc) {
  std::cout << "Error: transpose not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

// ---- Declared-only operations (not reached by bfs/sssp/pr/tc) -------------

template <typename T, typename a, typename b,
          typename SemiringT>
Info traceMxmTranspose(T* val, SemiringT op, const Matrix<a>* A, const Matrix<b>* B,
    Descriptor* desc) {
  std::cout << "Error: Trace operator not implemented!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename W, typename M, typename U, typename T>
Info scatter(Vector<W>* w, const Vector<M>* mask, const Vector<U>* u, T val,
    Descriptor* desc) {
  std::cout << "Error: scatter not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename W, typename U, typename M, typename I,
          typename BinaryOpT>
Info assignScatter(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    const Vector<U>* u, const Vector<I>* indices, Descriptor* desc) {
  std::cout << "Error: assignScatter not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename W, typename U, typename M, typename I,
          typename BinaryOpT>
Info extractGather(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    const Vector<U>* u, const Vector<I>* indices, Descriptor* desc) {
  std::cout << "Error: extractGather not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename W, typename a>
Info graphColor(Vector<W>* w, const Matrix<a>* A, Descriptor* desc) {
  std::cout << "Error: graphColor (cuSPARSE csrcolor) not implemented!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename W, typename U, typename a, typename M,
          typename BinaryOpT, typename SemiringT>
Info applyVxm(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Matrix<a>* A, Descriptor* desc) {
  std::cout << "Error: applyVxm not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}

template <typename c, typename a>
Info tril(Matrix<c>* C, Matrix<a>* A, Descriptor* desc) {
  Storage A_mat_type;
  CHECK(A->getStorage(&A_mat_type));

  if (reinterpret_cast<void*>(C) != reinterpret_cast<void*>(A))
    CHECK(C->dup(A));

  if (A_mat_type == GrB_SPARSE) {
    CHECK(C->setStorage(GrB_SPARSE));
    CHECK(trilSparse(&C->sparse_, &A->sparse_, desc));
  } else {
    std::cout << "Error: tril for dense matrix not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_OPERATIONS_HPP_
