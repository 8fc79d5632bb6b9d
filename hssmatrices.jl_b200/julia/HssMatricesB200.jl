# HssMatricesB200.jl — Julia binding of libhssb200.so (include/hssb200.h).
#
# Drop-in for ONE path of HssMatrices.jl v0.1.6: `hssA * X` / `mul!(C, hssA, X, α, β)` for
# HssMatrix{Float64} and strided Float64 matrices (src/matmul.jl:13-62).  The Julia API is
# unchanged: this file only adds MORE SPECIFIC methods of `*` and `mul!`; every other element
# type / array type keeps falling through to the reference methods.
#
# NOTE: Julia is not installed in the image this library was built in, so this file has never
# been executed there; the same C ABI is exercised by the Python ctypes harness
# (hssmatrices.jl_b200/__init__.py) that the tests and benchmarks use.  See INTEGRATION.md.
module HssMatricesB200

using HssMatrices
using HssMatrices: HssMatrix, isleaf, gensize
using LinearAlgebra
import Base: *, \
import LinearAlgebra: mul!

const libhssb = get(ENV, "HSSB200_LIB", joinpath(@__DIR__, "..", "lib", "libhssb200.so"))

struct HssbError <: Exception
  code::Cint
  msg::String
end

function check(rc::Integer)
  rc >= 0 && return rc
  msg = unsafe_string(ccall((:hssb_last_error, libhssb), Cstring, ()))
  rc == -2 && throw(DimensionMismatch(msg))   # same exception type as src/matmul.jl:19-20
  rc == -7 && throw(SingularException(0))     # what `D \ b` throws at src/ulvfactor.jl:83
  throw(HssbError(Cint(rc), msg))
end

"""Device-resident packed copy of an HssMatrix (hssb_matrix*)."""
mutable struct PackedHss
  handle::Ptr{Cvoid}
  m::Int
  n::Int
  function PackedHss(handle, m, n)
    p = new(handle, m, n)
    finalizer(p) do q
      q.handle == C_NULL || ccall((:hssb_destroy, libhssb), Cint, (Ptr{Cvoid},), q.handle)
      q.handle = C_NULL
    end
    return p
  end
end
Base.size(p::PackedHss) = (p.m, p.n)
Base.size(p::PackedHss, d::Integer) = size(p)[d]

dptr(A::Matrix{Float64}) = isempty(A) ? Ptr{Float64}(C_NULL) : pointer(A)

# Post-order walk of the pointer tree (src/hssmatrix.jl:11-34); the C++ packer copies every block
# during the call, flattens the tree into level-ordered arrays and uploads them.
function addnode(b::Ptr{Cvoid}, h::HssMatrix{Float64}, isroot::Bool)
  if isleaf(h)
    m, n = size(h.D)
    kr, kw = isroot ? (0, 0) : (size(h.U, 2), size(h.V, 2))
    GC.@preserve h begin
      return check(ccall((:hssb_builder_add_leaf, libhssb), Int64,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
        b, m, n, kr, kw, dptr(h.D), max(m, 1), dptr(h.U), max(m, 1), dptr(h.V), max(n, 1)))
    end
  end
  l = addnode(b, h.A11, false)
  r = addnode(b, h.A22, false)
  kr1, kw1 = gensize(h.A11); kr2, kw2 = gensize(h.A22)
  if isroot   # rooted(): src/hssmatrix.jl:266, used at src/matmul.jl:24
    GC.@preserve h begin
      return check(ccall((:hssb_builder_add_branch, libhssb), Int64,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
         Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
        b, l, r, 0, 0, dptr(h.B12), max(kr1, 1), dptr(h.B21), max(kr2, 1),
        C_NULL, 1, C_NULL, 1, C_NULL, 1, C_NULL, 1))
    end
  end
  kr, kw = gensize(h)
  GC.@preserve h begin
    return check(ccall((:hssb_builder_add_branch, libhssb), Int64,
      (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
       Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
      b, l, r, kr, kw, dptr(h.B12), max(kr1, 1), dptr(h.B21), max(kr2, 1),
      dptr(h.R1), max(kr1, 1), dptr(h.W1), max(kw1, 1), dptr(h.R2), max(kr2, 1), dptr(h.W2), max(kw2, 1)))
  end
end

"""
    pack(hssA; device=0) -> PackedHss

Flatten `hssA` (treated as root, like `rooted`) and upload it to GPU `device`.
"""
function pack(hssA::HssMatrix{Float64}; device::Integer=0)
  bref = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:hssb_builder_create, libhssb), Cint, (Ref{Ptr{Cvoid}},), bref))
  href = Ref{Ptr{Cvoid}}(C_NULL)
  try
    root = addnode(bref[], hssA, true)
    check(ccall((:hssb_builder_finalize, libhssb), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                bref[], root, device, 0, 1, href))
  finally
    ccall((:hssb_builder_destroy, libhssb), Cvoid, (Ptr{Cvoid},), bref[])
  end
  return PackedHss(href[], size(hssA, 1), size(hssA, 2))
end

"""
    pack(hssA, devices::AbstractVector{<:Integer}) -> PackedGroup

One process, one `ccall`, several GPUs (SURVEY §8b): the tree is registered ONCE and sharded by subtree over
`devices` (a power of two of them) inside the library (`hssb_group_finalize`); `*` / `mul!` on the result take the
whole X and Y and drive every device from the calling Julia thread.  No MPI.jl, no Distributed.jl.
"""
mutable struct PackedGroup
  handle::Ptr{Cvoid}
  m::Int
  n::Int
  function PackedGroup(handle, m, n)
    g = new(handle, m, n)
    finalizer(g) do q
      q.handle == C_NULL || ccall((:hssb_group_destroy, libhssb), Cint, (Ptr{Cvoid},), q.handle)
      q.handle = C_NULL
    end
    return g
  end
end
Base.size(g::PackedGroup) = (g.m, g.n)
Base.size(g::PackedGroup, d::Integer) = size(g)[d]

function pack(hssA::HssMatrix{Float64}, devices::AbstractVector{<:Integer})
  bref = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:hssb_builder_create, libhssb), Cint, (Ref{Ptr{Cvoid}},), bref))
  gref = Ref{Ptr{Cvoid}}(C_NULL)
  devs = Cint.(collect(devices))
  try
    root = addnode(bref[], hssA, true)
    check(ccall((:hssb_group_finalize, libhssb), Cint, (Ptr{Cvoid}, Int64, Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}),
                bref[], root, devs, length(devs), gref))
  finally
    ccall((:hssb_builder_destroy, libhssb), Cvoid, (Ptr{Cvoid},), bref[])
  end
  return PackedGroup(gref[], size(hssA, 1), size(hssA, 2))
end

function mul!(C::StridedMatrix{Float64}, g::PackedGroup, B::StridedMatrix{Float64}, α::Real, β::Real)
  size(g, 2) == size(B, 1) || throw(DimensionMismatch("First dimension of B does not match second dimension of A. Expected $(size(g, 2)), got $(size(B, 1))"))
  size(C) == (size(g, 1), size(B, 2)) || throw(DimensionMismatch("Dimensions of C don't match up with A and B."))
  (stride(B, 1) == 1 && stride(C, 1) == 1) || throw(ArgumentError("B and C need unit stride in the first dimension"))
  GC.@preserve B C begin
    check(ccall((:hssb_group_matmul, libhssb), Cint,
      (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64),
      g.handle, size(C, 1), size(B, 1), size(B, 2), pointer(B), max(stride(B, 2), 1), pointer(C), max(stride(C, 2), 1),
      Float64(α), Float64(β)))
  end
  return C
end
*(g::PackedGroup, B::StridedMatrix{Float64}) = mul!(Matrix{Float64}(undef, size(g, 1), size(B, 2)), g, B, 1.0, 0.0)   # src/matmul.jl:13
*(g::PackedGroup, x::StridedVector{Float64}) = reshape(g * reshape(x, length(x), 1), length(x))                      # src/matmul.jl:15

# mul!(C, hssA, B, α, β): src/matmul.jl:18-28.  β == 0 never reads C (src/matmul.jl:13 passes
# uninitialised memory).
function mul!(C::StridedMatrix{Float64}, p::PackedHss, B::StridedMatrix{Float64}, α::Real, β::Real)
  size(p, 2) == size(B, 1) || throw(DimensionMismatch("First dimension of B does not match second dimension of A. Expected $(size(p, 2)), got $(size(B, 1))"))
  size(C) == (size(p, 1), size(B, 2)) || throw(DimensionMismatch("Dimensions of C don't match up with A and B."))
  (stride(B, 1) == 1 && stride(C, 1) == 1) || throw(ArgumentError("B and C need unit stride in the first dimension"))
  GC.@preserve B C begin
    check(ccall((:hssb_matmul, libhssb), Cint,
      (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64),
      p.handle, size(C, 1), size(B, 1), size(B, 2), pointer(B), max(stride(B, 2), 1), pointer(C), max(stride(C, 2), 1),
      Float64(α), Float64(β)))
  end
  return C
end
*(p::PackedHss, B::StridedMatrix{Float64}) = mul!(Matrix{Float64}(undef, size(p, 1), size(B, 2)), p, B, 1.0, 0.0)   # src/matmul.jl:13
*(p::PackedHss, x::StridedVector{Float64}) = reshape(p * reshape(x, length(x), 1), length(x))                      # src/matmul.jl:15

# A * hssB (src/matmul.jl:14) without the adjoint copy of src/hssmatrix.jl:165-171:
# A*hssB = (hssB' * A')', and hssB' * X is hssb_matmul_t on the same packed generators.
function tmul!(C::StridedMatrix{Float64}, p::PackedHss, B::StridedMatrix{Float64}, α::Real, β::Real)
  size(p, 1) == size(B, 1) || throw(DimensionMismatch("First dimension of B does not match first dimension of A."))
  size(C) == (size(p, 2), size(B, 2)) || throw(DimensionMismatch("Dimensions of C don't match up with A' and B."))
  GC.@preserve B C begin
    check(ccall((:hssb_matmul_t, libhssb), Cint,
      (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64),
      p.handle, size(C, 1), size(B, 1), size(B, 2), pointer(B), max(stride(B, 2), 1), pointer(C), max(stride(C, 2), 1),
      Float64(α), Float64(β)))
  end
  return C
end
*(A::StridedMatrix{Float64}, p::PackedHss) = copy(tmul!(Matrix{Float64}(undef, size(p, 2), size(A, 1)), p, copy(A'), 1.0, 0.0)')

# hssA \ B (src/hssmatrix.jl:234) = ulvfactsolve (src/ulvfactor.jl:10-19).  The reference factorises and
# solves in one pass on every call; the library factorises once per packed handle (hssb_ulv_factor, on
# the device) and every further solve only applies the stored factors.
function HssMatrices.ulvfactsolve(p::PackedHss, B::StridedMatrix{Float64})
  size(p, 1) == size(B, 1) || throw(DimensionMismatch("First dimension of B does not match first dimension of A."))
  stride(B, 1) == 1 || throw(ArgumentError("B needs unit stride in the first dimension"))
  Z = Matrix{Float64}(undef, size(p, 2), size(B, 2))
  GC.@preserve B Z begin
    check(ccall((:hssb_solve, libhssb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
      p.handle, size(B, 1), size(B, 2), pointer(B), max(stride(B, 2), 1), pointer(Z), max(size(Z, 1), 1)))
  end
  return Z
end
\(p::PackedHss, B::StridedMatrix{Float64}) = HssMatrices.ulvfactsolve(p, B)
\(p::PackedHss, b::StridedVector{Float64}) = reshape(HssMatrices.ulvfactsolve(p, reshape(b, length(b), 1)), length(b))
# A / hssB (src/hssmatrix.jl:236: ulvfactsolve(hssB', collect(A'))') without building hssB': uniform trees only
# (second factor pool from the adjoint twin pool); other trees throw HssbError -- pack(hssB') and use `\` there.
function Base.:/(A::StridedMatrix{Float64}, p::PackedHss)
  size(A, 2) == size(p, 1) || throw(DimensionMismatch("Second dimension of A does not match first dimension of hssB."))
  Bt = copy(A')
  Z = Matrix{Float64}(undef, size(p, 1), size(Bt, 2))
  GC.@preserve Bt Z begin
    check(ccall((:hssb_solve_t, libhssb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
      p.handle, size(Bt, 1), size(Bt, 2), pointer(Bt), max(size(Bt, 1), 1), pointer(Z), max(size(Z, 1), 1)))
  end
  return copy(Z')
end
"""Factorise ahead of time (otherwise the first `\\` does it)."""
ulvfactor!(p::PackedHss) = (check(ccall((:hssb_ulv_factor, libhssb), Cint, (Ptr{Cvoid},), p.handle)); p)

# ---- drop-in methods on HssMatrix{Float64} ---------------------------------------------------
# HssMatrix is mutable (recompress!, prune_leaves!, field assignment as in test/runtests.jl:75), so a
# device copy cached per object can go stale.  The cache therefore
#   * holds its keys WEAKLY (WeakKeyDict): an HssMatrix the caller drops is collected as usual, and the
#     finalizer of its PackedHss then frees the device pool;
#   * stores a fingerprint of the generator tree next to the handle (per generator: data pointer, size
#     and three sampled entries, O(#nodes), no pass over the data) and re-packs when it changed.  This
#     catches field assignment, recompress!, prune_leaves! and orthonormalize_generators! (they replace
#     or resize the generator arrays); an in-place edit of single entries that misses the three samples
#     needs invalidate!(hssA).
# Only `*` and `mul!` (src/matmul.jl:13-28) are overridden for HssMatrix.  The solver entries
# (`\`, ulvfactsolve) are deliberately NOT: the reference routes `/(A, hssB)` through
# `ulvfactsolve(hssB', ...)` on a temporary adjoint (src/hssmatrix.jl:236), which would pack, factorise
# and cache a device copy of an object the caller never sees.  Use `pack(hssA) \ B` for the GPU solver.
struct CacheEntry
  packed::PackedHss
  fingerprint::UInt64
end
const CACHE = WeakKeyDict{HssMatrix{Float64}, CacheEntry}()
const CACHE_LOCK = ReentrantLock()

@inline function fp_block(h::UInt64, A::Matrix{Float64})
  h = hash(UInt(pointer(A)), hash(size(A), h))
  n = length(A)
  n == 0 && return h
  @inbounds return hash(A[1], hash(A[(n + 1) >> 1], hash(A[n], h)))
end
function fingerprint(hssA::HssMatrix{Float64}, isroot::Bool=true, h::UInt64=UInt64(0x48535342))
  if isleaf(hssA)
    h = fp_block(h, hssA.D)
    isroot || (h = fp_block(fp_block(h, hssA.U), hssA.V))
    return h
  end
  h = fingerprint(hssA.A11, false, h)
  h = fingerprint(hssA.A22, false, h)
  h = fp_block(fp_block(h, hssA.B12), hssA.B21)
  isroot || (h = fp_block(fp_block(fp_block(fp_block(h, hssA.R1), hssA.W1), hssA.R2), hssA.W2))
  return h
end

function packed(hssA::HssMatrix{Float64})
  fp = fingerprint(hssA)
  lock(CACHE_LOCK) do
    e = get(CACHE, hssA, nothing)
    if e === nothing || e.fingerprint != fp
      e = CacheEntry(pack(hssA), fp)   # the replaced PackedHss is freed by its finalizer
      CACHE[hssA] = e
    end
    return e.packed
  end
end
"""Forget the device copy of `hssA` (only needed after editing single entries of a generator in place)."""
invalidate!(hssA::HssMatrix{Float64}) = (lock(() -> delete!(CACHE, hssA), CACHE_LOCK); nothing)

mul!(C::StridedMatrix{Float64}, hssA::HssMatrix{Float64}, B::StridedMatrix{Float64}, α::Real, β::Real) = mul!(C, packed(hssA), B, α, β)
*(hssA::HssMatrix{Float64}, B::StridedMatrix{Float64}) = packed(hssA) * B
*(A::StridedMatrix{Float64}, hssB::HssMatrix{Float64}) = A * packed(hssB)

export pack, PackedHss, PackedGroup, invalidate!, ulvfactor!

end # module
