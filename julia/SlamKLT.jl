# SlamKLT.jl -- drop-in shim that forwards SLAM.jl's KLT front-end methods to libslamklt.so via ccall.
#
# COMPILE-UNTESTED: Julia is not installed in the image this library was built in.  The file is the binding a
# SLAM.jl maintainer would `include` from src/SLAM.jl *instead of* optical_flow/{pyramid,lucas_kanade,utils}.jl,
# tracker.jl and extractor.jl.  It names none of the types SLAM.jl defines further down its include list (Camera, Frame,
# MapManager): the two optional whole-method replacements that do live in julia/SlamKLTOptional.jl, included after mapper.jl.
# Call sites (front_end.jl:459-467, map_manager.jl:104-105,517-521,549-551, mapper.jl:51-60, SLAM.jl:158-160,216-219) stay unchanged.
#
# Memory-layout facts used (SLAM.jl:22-26): Vector{SVector{2,Float64}} is a dense N x 2 Float64 array in (y, x)
# order; Matrix{Gray{Float64}} reinterprets to column-major Float64 with ld = H; Vector{CartesianIndex{2}} is dense
# Int64 pairs.

import Random   # BRIEF's sampling pattern is drawn from Julia's RNG (see _brief_pairs); BRIEF itself comes from ImageFeatures,
                # which src/SLAM.jl already loads

const libslamklt = get(ENV, "SLAMKLT_LIB", "libslamklt.so")

const SLAMKLT_F64 = Cint(0)
const SLAMKLT_MODE_UPDATE = Cint(0)
const SLAMKLT_MODE_CTOR = Cint(1)

struct SlamKltLKParams
    iterations::Int32
    window_size::Int32
    pyramid_levels::Int32
    reserved::Int32
    eigenvalue_threshold::Float64
    epsilon::Float64
    max_distance::Float64
end

struct SlamKltDetectParams
    max_points::Int32
    radius::Int32
    grid_h::Int32
    grid_w::Int32
    cell_size::Int32
    reserved::Int32
    sigma_mask::Float64
    min_response::Float64
end

@inline function _ck(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:slamklt_last_error, libslamklt), Cstring, ()))
    # optflow! throws a String when the pyramids are too shallow (lucas_kanade.jl:12-15)
    rc == -3 ? throw("Not enough layers in pyramids.") : error("[slamklt $rc] $msg")
end

# ---- context: one per Julia process and device; the library serialises concurrent calls (front-end + mapper tasks)
mutable struct KltContext
    handle::Ptr{Cvoid}
    function KltContext(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        _ck(ccall((:slamklt_ctx_create, libslamklt), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        c = new(h[])
        finalizer(c -> ccall((:slamklt_ctx_destroy, libslamklt), Cint, (Ptr{Cvoid},), c.handle), c)
        c
    end
end
const KLT_CTX = Ref{KltContext}()
klt_ctx() = (isassigned(KLT_CTX) || (KLT_CTX[] = KltContext(parse(Int, get(ENV, "SLAMKLT_DEVICE", "0")))); KLT_CTX[])

# ---- LKPyramid: same type name and parameters {G, C} as pyramid.jl:16-24 so that front_end.jl:32-40 and mapper.jl:3 compile.
struct LKCache end
mutable struct LKPyramid{G, C}
    handle::Ptr{Cvoid}     # slamklt_pyr*, C_NULL for the gradient-less empty pyramid of front_end.jl:38-40
    size::Tuple{Int, Int}
    levels::Int
end
has_cache(::LKPyramid{G, C}) where {G, C} = C !== Nothing
has_gradients(::LKPyramid{G, C}) where {G, C} = G !== Nothing

# the 7-positional-argument literal used for empty pyramids (front_end.jl:38-40, 507-509)
LKPyramid(layers::Vector, ::Nothing, ::Nothing, ::Nothing, ::Nothing, ::Nothing, ::Nothing) =
    LKPyramid{Nothing, Nothing}(C_NULL, (0, 0), 0)

function _new_pyramid(H, W, levels)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _ck(ccall((:slamklt_pyr_create, libslamklt), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
              klt_ctx().handle, H, W, levels, h))
    p = LKPyramid{Vector{Matrix{Gray{Float64}}}, LKCache}(h[], (H, W), levels)
    finalizer(p -> ccall((:slamklt_pyr_destroy, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), klt_ctx().handle, p.handle), p)
    p
end

function _build!(p::LKPyramid, image::AbstractMatrix, σ, mode)
    img = reinterpret(Float64, image)           # Matrix{Gray{Float64}} -> Matrix{Float64}, no copy
    GC.@preserve img _ck(ccall((:slamklt_pyr_build, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cdouble, Cint),
        klt_ctx().handle, p.handle, pointer(img), SLAMKLT_F64, size(img, 1), σ, mode))
    p
end

# pyramid.jl:40-79
function LKPyramid(image, levels; downsample = 2, σ = 1.0, gradients = true, reusable = false)
    p = _new_pyramid(size(image, 1), size(image, 2), levels)
    _build!(p, image, σ, SLAMKLT_MODE_CTOR)
end
# pyramid.jl:81-96
update!(lk::LKPyramid, img; σ = 1.0) = _build!(lk, img, σ, SLAMKLT_MODE_UPDATE)
# pyramid.jl:28-38
function Base.copy!(dst::LKPyramid, src::LKPyramid)
    _ck(ccall((:slamklt_pyr_copy, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), klt_ctx().handle, dst.handle, src.handle))
    dst
end
# SLAM.jl:216-219 hands deepcopy(current_pyramid) to the mapper task
function Base.deepcopy_internal(p::LKPyramid{G, C}, ::IdDict) where {G, C}
    p.handle == C_NULL && return LKPyramid{G, C}(C_NULL, p.size, p.levels)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _ck(ccall((:slamklt_pyr_clone, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), klt_ctx().handle, p.handle, h))
    q = LKPyramid{G, C}(h[], p.size, p.levels)
    finalizer(q -> ccall((:slamklt_pyr_destroy, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), klt_ctx().handle, q.handle), q)
    q
end

# ---- lucas_kanade.jl:1-7
Base.@kwdef struct LucasKanade
    iterations::Int64 = 30
    window_size::Int64 = 9
    pyramid_levels::Int64 = 3
    eigenvalue_threshold::Float64 = 1e-4
    ϵ::Float64 = 1e-2
end
_params(a::LucasKanade, max_distance) = SlamKltLKParams(a.iterations, a.window_size, a.pyramid_levels, 0,
    a.eigenvalue_threshold, a.ϵ, max_distance)

# lucas_kanade.jl:9-100 (displacement is mutated in place like the reference)
function optflow!(displacement::Vector{Point2f}, first_pyramid::LKPyramid, second_pyramid::LKPyramid,
                  points::Vector{Point2f}, algorithm::LucasKanade)
    n = length(points)
    status = Vector{UInt8}(undef, n)
    n_good = Ref{Cint}(0)
    prm = Ref(_params(algorithm, 0.0))
    GC.@preserve displacement points status _ck(ccall((:slamklt_optflow, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ref{SlamKltLKParams}, Ptr{UInt8}, Ref{Cint}),
        klt_ctx().handle, first_pyramid.handle, second_pyramid.handle, pointer(points), pointer(displacement), n, prm, status, n_good))
    displacement, status .!= 0, Int(n_good[])
end

# tracker.jl:17-68; `keypoints` / `displacement` may be SubArray views (map_manager.jl:511-513) => collect to dense
function fb_tracking!(new_keypoints::AbstractVector{Point2f}, previous_pyramid::LKPyramid, current_pyramid::LKPyramid,
                      keypoints::AbstractVector{Point2f}, algorithm::LucasKanade;
                      displacement::Union{Nothing, AbstractVector{Point2f}} = nothing, max_distance::Real = 0.5)
    isempty(keypoints) && return
    kps = keypoints isa Vector ? keypoints : collect(keypoints)
    disp = displacement === nothing ? nothing : (displacement isa Vector ? displacement : collect(displacement))
    out = new_keypoints isa Vector ? new_keypoints : Vector{Point2f}(undef, length(kps))
    n = length(kps)
    status = Vector{UInt8}(undef, n)
    prm = Ref(_params(algorithm, Float64(max_distance)))
    GC.@preserve kps disp out status _ck(ccall((:slamklt_fb_track, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ref{SlamKltLKParams}, Ptr{Float64}, Ptr{UInt8}),
        klt_ctx().handle, previous_pyramid.handle, current_pyramid.handle, pointer(kps),
        disp === nothing ? Ptr{Float64}(C_NULL) : pointer(disp), n, prm, pointer(out), status))
    out === new_keypoints || copyto!(new_keypoints, out)
    new_keypoints, (status .& 0x01) .!= 0     # any AbstractVector{Bool} works for the callers (status[j] only)
end

function fb_tracking!(previous_pyramid::LKPyramid, current_pyramid::LKPyramid, keypoints::AbstractVector{Point2f};
                      displacement = nothing, iterations::Int = 30, window_size::Int = 11, pyramid_levels::Int = 3,
                      max_distance::Real = 0.5)
    new_keypoints = Vector{Point2f}(undef, length(keypoints))
    algorithm = LucasKanade(; iterations, window_size, pyramid_levels)
    fb_tracking!(new_keypoints, previous_pyramid, current_pyramid, keypoints, algorithm; displacement, max_distance)
end

# ---- extractor.jl:7-22.  The descriptor field stays (params.do_local_matching = true needs `describe`).
struct Extractor
    max_points::Int
    descriptor::BRIEF
    radius::Int
    grid_resolution::Tuple{Int, Int}
    cell_size::Int
end
Extractor(max_points, radius, grid_resolution, cell_size) =
    Extractor(max_points, BRIEF(; size = 256), radius, grid_resolution, cell_size)       # extractor.jl:20-22

# BRIEF's sampling pattern comes from Julia's RNG (ImageFeatures brief.jl: Random.seed!(params.seed); params.sampling_type(size,
# window)); it is drawn here exactly as create_descriptor draws it and handed to the library as (dy1, dx1, dy2, dx2) rows.
function _brief_pairs(b::BRIEF)
    Random.seed!(b.seed)                                   # ImageFeatures brief.jl does the same before sampling
    s1, s2 = b.sampling_type(b.size, b.window)
    pairs = Matrix{Int32}(undef, 4, b.size)
    for k in 1:b.size
        pairs[1, k] = s1[k][1]; pairs[2, k] = s1[k][2]; pairs[3, k] = s2[k][1]; pairs[4, k] = s2[k][2]
    end
    pairs
end

# extractor.jl:103-105: create_descriptor(image, keypoints, e.descriptor) -> (Vector{BitVector}, Vector{CartesianIndex{2}})
function describe(e::Extractor, image, keypoints)
    b = e.descriptor
    img = reinterpret(Float64, image)
    H, W = size(img)
    n = length(keypoints)
    n == 0 && return BitVector[], CartesianIndex{2}[]
    kps = keypoints isa Vector{CartesianIndex{2}} ? keypoints : collect(CartesianIndex{2}, keypoints)
    pairs = _brief_pairs(b)
    words = b.size ÷ 32
    desc = Matrix{UInt32}(undef, words, n)
    valid = Vector{UInt8}(undef, n)
    GC.@preserve img kps pairs desc valid _ck(ccall((:slamklt_describe, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Int64}, Cint, Ptr{Int32}, Cint, Cint, Cdouble, Ptr{UInt32}, Ptr{UInt8}),
        klt_ctx().handle, pointer(img), SLAMKLT_F64, H, W, H, pointer(kps), n, pointer(pairs), b.size, b.window, b.sigma,
        pointer(desc), pointer(valid)))
    descriptors = BitVector[]; kept = CartesianIndex{2}[]
    for j in 1:n
        valid[j] == 0 && continue
        d = BitVector(undef, b.size)
        for w in 1:words, bit in 0:31                      # bit b of word w = descriptor bit 32 (w - 1) + b + 1
            d[32 * (w - 1) + bit + 1] = (desc[w, j] >> bit) & 0x1 == 0x1
        end
        push!(descriptors, d); push!(kept, kps[j])
    end
    descriptors, kept
end

# extractor.jl:63-95
function detect(e::Extractor, image, current_points; σ_mask = 3)
    length(current_points) ≥ e.max_points && return CartesianIndex{2}[]
    img = reinterpret(Float64, image)
    H, W = size(img)
    cur = current_points isa Vector{Point2f} ? current_points : collect(Point2f, current_points)
    n_cells = e.grid_resolution[1] * e.grid_resolution[2]
    cap = max(1, min(cld(e.max_points - length(cur), n_cells), e.cell_size^2) * n_cells)
    out = Vector{CartesianIndex{2}}(undef, cap)
    n_out = Ref{Cint}(0)
    prm = Ref(SlamKltDetectParams(e.max_points, e.radius, e.grid_resolution[1], e.grid_resolution[2], e.cell_size, 0,
                                  Float64(σ_mask), 1e-4))
    GC.@preserve img cur out _ck(ccall((:slamklt_detect, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Float64}, Cint, Ref{SlamKltDetectParams}, Ptr{Int64}, Cint, Ref{Cint}),
        klt_ctx().handle, pointer(img), SLAMKLT_F64, H, W, H, isempty(cur) ? Ptr{Float64}(C_NULL) : pointer(cur), length(cur),
        prm, pointer(out), cap, n_out))
    resize!(out, n_out[])
end

# ---- optional: many frames of a stream per call (INTEGRATION.md §4).  Not used by SLAM.jl's own call sites; a consumer that
# follows several independent streams keeps one KltStreamBatch per stream and alternates step_begin! / step_end! so that the
# uploads of one step hide behind the kernels of the other.
mutable struct KltStreamBatch
    handle::Ptr{Cvoid}
    H::Int; W::Int; n_frames::Int; max_points::Int
    keep::Any    # buffers of the step in flight (they must stay alive and untouched until step_end!)
    function KltStreamBatch(H::Integer, W::Integer, levels::Integer, n_frames::Integer, max_points::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        _ck(ccall((:slamklt_batch_create, libslamklt), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                  klt_ctx().handle, H, W, levels, n_frames, max_points, h))
        b = new(h[], H, W, n_frames, max_points, nothing)
        finalizer(b -> ccall((:slamklt_batch_destroy, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), klt_ctx().handle, b.handle), b)
        b
    end
end

"first frame of the stream (slot 0 of the batch)"
function prime!(b::KltStreamBatch, image::AbstractMatrix; σ = 1.0)
    img = reinterpret(Float64, image)
    GC.@preserve img _ck(ccall((:slamklt_batch_prime, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cdouble, Cint),
                               klt_ctx().handle, b.handle, pointer(img), SLAMKLT_F64, b.H, σ, SLAMKLT_MODE_UPDATE))
    b
end

"""
    step_begin!(batch, frames, keypoints, alg; max_distance, σ)

`frames`: H × W × n_frames array of `Gray{Float64}` / `Float64` (the next n_frames images of the stream), `keypoints`: n × n_frames
matrix of `Point2f` (the points to track from frame i-1 to frame i).  Queues upload, pyramid construction (`update!`), forward-backward
tracking and the result copies; returns at once.  `step_end!` returns `(new_keypoints, status)` like `fb_tracking!` does per frame.
"""
function step_begin!(b::KltStreamBatch, frames::AbstractArray{<:Any, 3}, keypoints::AbstractMatrix{Point2f}, alg::LucasKanade;
                     max_distance = 1.0, σ = 1.0)
    img = reinterpret(Float64, frames)
    n = size(keypoints, 1)
    new_keypoints = similar(keypoints)
    status = Vector{UInt8}(undef, n * b.n_frames)
    prm = Ref(SlamKltLKParams(alg.iterations, alg.window_size, alg.pyramid_levels, 0, alg.eigenvalue_threshold, alg.ϵ, max_distance))
    b.keep = (img, keypoints, new_keypoints, status, prm)
    _ck(ccall((:slamklt_batch_step_begin, libslamklt), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Csize_t, Ptr{Float64}, Cint, Cdouble, Cint, Ref{SlamKltLKParams}, Ptr{Float64},
               Ptr{UInt8}),
              klt_ctx().handle, b.handle, pointer(img), SLAMKLT_F64, b.H, b.H * b.W * 8, pointer(keypoints), n, σ, SLAMKLT_MODE_UPDATE, prm,
              pointer(new_keypoints), pointer(status)))
    b
end

function step_end!(b::KltStreamBatch)
    _ck(ccall((:slamklt_batch_step_end, libslamklt), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), klt_ctx().handle, b.handle))
    _, keypoints, new_keypoints, status, _ = b.keep
    b.keep = nothing
    new_keypoints, reshape((status .& 0x01) .!= 0, size(keypoints))
end

"one synchronous step"
step!(b::KltStreamBatch, frames, keypoints, alg::LucasKanade; kw...) = (step_begin!(b, frames, keypoints, alg; kw...); step_end!(b))
