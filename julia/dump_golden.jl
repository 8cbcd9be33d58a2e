# dump_golden.jl -- pins the CPU oracle (oracle/) against the real SLAM.jl.
#
# COMPILE-UNTESTED: there is no Julia in the image this repository is built in (SURVEY 8c).  Run it wherever Julia + SLAM.jl
# are installed:
#
#     python tools/golden_io.py export /tmp/klt_golden          # writes the inputs (raw little-endian Float64, column-major)
#     julia --project=/path/to/SLAM.jl julia/dump_golden.jl /tmp/klt_golden
#     python tools/golden_io.py check /tmp/klt_golden           # compares every dumped array with the oracle
#
# File format: <name>.bin = Int64 ndims, Int64 dims..., then the Float64 (or Int64 for *_i64) payload in Julia's own
# (column-major) order.  Everything the oracle restates is dumped: pyramid layers, Scharr gradients, integral images of the
# smoothed gradient products (both border regimes: constructor and update!), optflow! / fb_tracking! results, detect output,
# and the camera geometry used by optical_flow_matching!.
using SLAM
using SLAM: LKPyramid, update!, LucasKanade, optflow!, fb_tracking!, Extractor, detect, Camera, Point2f, Point3f
using SLAM: undistort_point, backproject, project_undistort, in_image
using Images, StaticArrays

dir = ARGS[1]

function rd(name)
    open(joinpath(dir, name * ".bin")) do io
        nd = read(io, Int64)
        dims = ntuple(_ -> Int(read(io, Int64)), nd)
        a = Array{Float64}(undef, dims...)
        read!(io, a)
        a
    end
end
function wr(name, a::AbstractArray{T}) where T <: Union{Float64, Int64}
    open(joinpath(dir, name * ".bin"), "w") do io
        write(io, Int64(ndims(a)))
        for d in size(a) write(io, Int64(d)) end
        write(io, collect(a))
    end
end
wrimg(name, m) = wr(name, Float64.(m))
points(a) = [Point2f(a[1, i], a[2, i]) for i in 1:size(a, 2)]           # stored as 2 x N, (y, x) per column
unpoints(v) = reduce(hcat, [Float64[p[1], p[2]] for p in v]; init = zeros(2, 0))

img0 = Gray{Float64}.(rd("img0")); img1 = Gray{Float64}.(rd("img1"))
pts = points(rd("pts"))
meta = rd("meta")                                                       # levels, window, max_distance, max_points, radius, grid_h, grid_w, cell
levels, window = Int(meta[1]), Int(meta[2]); max_distance = meta[3]

# ---- pyramids: constructor regime (pyramid.jl:40-79) and update! regime (pyramid.jl:81-96)
p0 = LKPyramid(img0, levels; reusable = true)
p1 = LKPyramid(img1, levels; reusable = true)
for (tag, p) in (("ctor0", p0), ("ctor1", p1))
    for l in 1:levels + 1
        wrimg("$(tag)_layer$(l - 1)", p.layers[l]); wrimg("$(tag)_Iy$(l - 1)", p.Iy[l]); wrimg("$(tag)_Ix$(l - 1)", p.Ix[l])
        wrimg("$(tag)_Iyy$(l - 1)", p.Iyy[l]); wrimg("$(tag)_Ixx$(l - 1)", p.Ixx[l]); wrimg("$(tag)_Iyx$(l - 1)", p.Iyx[l])
    end
end
update!(p1, img1)
for l in 1:levels + 1
    wrimg("upd1_layer$(l - 1)", p1.layers[l]); wrimg("upd1_Iy$(l - 1)", p1.Iy[l]); wrimg("upd1_Ix$(l - 1)", p1.Ix[l])
    wrimg("upd1_Iyy$(l - 1)", p1.Iyy[l]); wrimg("upd1_Ixx$(l - 1)", p1.Ixx[l]); wrimg("upd1_Iyx$(l - 1)", p1.Iyx[l])
end

# ---- optflow! (lucas_kanade.jl:9-100) from p0 (ctor) to p1 (updated), zero initial displacement
alg = LucasKanade(; iterations = 30, window_size = window, pyramid_levels = levels)
disp = fill(Point2f(0.0, 0.0), length(pts))
_, status, n_good = optflow!(disp, p0, p1, pts, alg)
wr("optflow_disp", unpoints(disp)); wr("optflow_status_i64", Int64.(collect(status)))

# ---- fb_tracking! (tracker.jl:17-82)
new_kps, fb_status = fb_tracking!(p0, p1, pts; pyramid_levels = levels, window_size = window, max_distance)
st = Int64.(collect(fb_status))
out = unpoints([st[i] == 1 ? new_kps[i] : Point2f(NaN, NaN) for i in eachindex(st)])
wr("fb_new", out); wr("fb_status_i64", st)

# ---- detect (extractor.jl:63-95)
e = Extractor(Int(meta[4]), Int(meta[5]), (Int(meta[6]), Int(meta[7])), Int(meta[8]))
kp = detect(e, img0, pts[1:10])
wr("detect_i64", reduce(hcat, [Int64[c[1], c[2]] for c in kp]; init = zeros(Int64, 2, 0)))
kp_all = detect(e, img0, Point2f[])
wr("detect_nomask_i64", reduce(hcat, [Int64[c[1], c[2]] for c in kp_all]; init = zeros(Int64, 2, 0)))

# ---- camera geometry (camera.jl:60-143) used by optical_flow_matching! (map_manager.jl:451-564)
c = rd("camera")                                                        # fx fy cx cy k1 k2 p1 p2 height width
cam = Camera(c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], Int(c[9]), Int(c[10]))
cw = SMatrix{4, 4, Float64, 16}(rd("cw"))
world = rd("world")                                                     # 3 x N
proj = [project_undistort(cam, cw * SVector{4, Float64}(world[1, i], world[2, i], world[3, i], 1.0)) for i in 1:size(world, 2)]
wr("cam_proj", unpoints(proj))
und = [undistort_point(cam, p) for p in pts]
wr("cam_undist", unpoints(und))
wr("cam_backproject", reduce(hcat, [collect(backproject(cam, u)) for u in und]))
println("dumped golden vectors into ", dir)
