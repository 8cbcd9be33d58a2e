# SlamKLTOptional.jl -- the optional device calls that replace whole SLAM.jl methods (SURVEY 8f rows 1-2):
# triangulate_stereo! (mapper.jl:142-183) and optical_flow_matching! (map_manager.jl:451-564).
#
# COMPILE-UNTESTED (no Julia in the image this library was built in).  These methods name SLAM.jl's own types (MapManager, Frame,
# Camera), so this file is included at the END of src/SLAM.jl's include list -- after mapper.jl -- where its methods replace the
# originals of the same signature; julia/SlamKLT.jl (included near the top, in place of the optical-flow files) must already be
# loaded.  Leaving this file out keeps SLAM.jl's own Julia bodies, which then call the shim's fb_tracking!.

struct SlamKltCamera
    fx::Float64; fy::Float64; cx::Float64; cy::Float64
    k1::Float64; k2::Float64; p1::Float64; p2::Float64
    height::Int64; width::Int64
    Ti0::NTuple{16, Float64}
end
_cam(c::Camera) = SlamKltCamera(c.fx, c.fy, c.cx, c.cy, c.k1, c.k2, c.p1, c.p2, c.height, c.width, Tuple(c.Ti0))

struct SlamKltMatchingParams
    lk::SlamKltLKParams
    stereo::Int32
    pyramid_levels_3d::Int32
    epipolar_error::Float64
end

# ---- optional (SURVEY 8f row 2): triangulate_stereo! (mapper.jl:142-183) with the per-keypoint DLT + checks on the device ------
# The GEEV4x4Cache argument is kept for the call site (mapper.jl:72-74) and ignored.
function triangulate_stereo!(map_manager::MapManager, frame::Frame, max_error, cache)
    stereo_keypoints = get_stereo_keypoints(frame)
    isempty(stereo_keypoints) && (@warn "[MP] No stereo keypoints to triangulate."; return)
    ids = Int64[]; und = Point2f[]; rund = Point2f[]
    for kp in stereo_keypoints                                     # host-side bookkeeping of mapper.jl:156-161
        kp.is_3d && continue
        mp = get_mappoint(map_manager, kp.id)
        mp ≡ nothing && (remove_mappoint_obs!(map_manager, kp.id, frame.kfid); continue)
        mp.is_3d && continue
        push!(ids, kp.id); push!(und, kp.undistorted_pixel); push!(rund, kp.right_undistorted_pixel)
    end
    n = length(ids)
    n == 0 && return
    world = Vector{Point3f}(undef, n); status = Vector{UInt8}(undef, n)
    cam = Ref(_cam(frame.camera)); rcam = Ref(_cam(frame.right_camera))
    wc = collect(Float64, lock(() -> frame.wc, frame.pose_lock))
    GC.@preserve und rund world status wc _ck(ccall((:slamklt_triangulate_stereo, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ref{SlamKltCamera}, Ref{SlamKltCamera}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{UInt8}),
        klt_ctx().handle, pointer(und), pointer(rund), n, cam, rcam, pointer(wc), Float64(max_error), pointer(world), pointer(status)))
    for j in 1:n
        status[j] == 0x01 ? update_mappoint!(map_manager, ids[j], world[j]) : remove_stereo_keypoint!(frame, ids[j])
    end
end

# ---- optional (SURVEY 8f rows 1-2): optical_flow_matching! with its per-keypoint geometry on the device ------------------
# Replaces the body of map_manager.jl:451-564: one ccall does the projection of the 3-D keypoints' map points, the prior
# displacement, both fb_tracking! passes, and update_keypoint! / maybe_stereo_update! arithmetic; the dictionary
# bookkeeping below is the part that has to stay in Julia.  Include this file *after* map_manager.jl to let this method
# replace the original, or call it under another name.
function optical_flow_matching!(map_manager::MapManager, frame, from_pyramid::LKPyramid, to_pyramid::LKPyramid, stereo)
    p = map_manager.params
    keypoints = stereo ? get_keypoints(frame) : collect(values(frame.keypoints))
    ids = Int64[]; pixels = Point2f[]; undist = Point2f[]; world = Point3f[]; is_3d = UInt8[]
    for kp in keypoints
        mp = nothing
        if kp.is_3d
            mp = stereo ? get_mappoint(map_manager, kp.id) : get(map_manager.map_points, kp.id, nothing)
            mp ≡ nothing && (remove_mappoint_obs!(map_manager, kp.id, frame.kfid); continue)   # map_manager.jl:479-482
        end
        push!(ids, kp.id); push!(pixels, kp.pixel); push!(undist, kp.undistorted_pixel)
        push!(is_3d, kp.is_3d ? 0x01 : 0x00)
        push!(world, kp.is_3d ? get_position(mp) : Point3f(0, 0, 0))
    end
    n = length(ids)
    n == 0 && return nothing
    out_pixel = Vector{Point2f}(undef, n); out_undist = Vector{Point2f}(undef, n); out_position = Vector{Point3f}(undef, n)
    status = Vector{UInt8}(undef, n)
    cw = lock(() -> frame.cw, frame.pose_lock)
    cam = Ref(_cam(frame.camera)); rcam = Ref(_cam(frame.right_camera))
    prm = Ref(SlamKltMatchingParams(
        SlamKltLKParams(30, p.window_size, p.pyramid_levels, 0, 1e-4, 1e-2, p.max_ktl_distance), stereo ? 1 : 0, 1, 2.0))
    cwv = collect(Float64, cw)     # column-major 16 values
    GC.@preserve pixels undist world is_3d out_pixel out_undist out_position status cwv _ck(ccall(
        (:slamklt_optical_flow_matching, libslamklt), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64},
         Ref{SlamKltCamera}, Ref{SlamKltCamera}, Ref{SlamKltMatchingParams}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
        klt_ctx().handle, from_pyramid.handle, to_pyramid.handle, pointer(pixels), pointer(is_3d), pointer(world),
        pointer(undist), n, pointer(cwv), cam, rcam, prm, pointer(out_pixel), pointer(out_undist), pointer(out_position),
        pointer(status)))
    for j in 1:n
        s = status[j]
        if s & 0x01 != 0
            lock(frame.keypoints_lock) do
                ckp = get(frame.keypoints, ids[j], nothing)
                ckp ≡ nothing && return
                if stereo                                    # update_stereo_keypoint!, frame.jl:272-287
                    ckp.right_pixel = out_pixel[j]
                    ckp.right_undistorted_pixel = out_undist[j]
                    ckp.right_position = out_position[j]
                    ckp.is_stereo || (ckp.is_stereo = true; frame.nb_stereo_kpts += 1)
                else                                         # update_keypoint!, frame.jl:252-270
                    kp = deepcopy(ckp)
                    kp.pixel = out_pixel[j]; kp.undistorted_pixel = out_undist[j]; kp.position = out_position[j]
                    kp.is_stereo && (kp.is_stereo = false; frame.nb_stereo_kpts -= 1)
                    update_keypoint_in_grid!(frame, ckp, kp)
                    frame.keypoints[ids[j]] = kp
                end
            end
        elseif s & 0x08 != 0
            stereo && remove_mappoint_obs!(map_manager, ids[j], frame.kfid)           # map_manager.jl:496-498
        elseif !stereo
            remove_obs_from_current_frame!(map_manager, ids[j])                       # map_manager.jl:558
        end
    end
    nothing
end

