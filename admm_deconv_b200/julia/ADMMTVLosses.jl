# ADMMTVLosses.jl -- Julia drop-in for /root/reference/src/metrics/gmsd.jl (+ iqa_utils.jl) and
# src/metrics/ssim.jl over the C ABI of include/admmtv_loss.h (libadmmtv.so, sm_100a kernels).
#
# STATUS: written to the ABI, NOT EXECUTED (no Julia toolchain in the build image or on the GPU box); the
# identical ABI is exercised from Python (admm_deconv_b200/losses.py, tests/test_gpu_losses.py).  Usage:
#
#     # src/train.jl, src/train_v2.jl:
#     # replace  include("metrics/gmsd.jl"); include("metrics/ssim.jl")  by  include("ADMMTVLosses.jl"); using .ADMMTVLosses
#
# Same names and argument order: gmsd(x, y, t, α, reduction), gmsd_loss, ssim(x, y, kernel_ref; peakval, crop, dims),
# ssim_loss, ssim_loss_fast, ssim_kernel.
# The reduction happens on the device; the value is read back as a Float32 scalar (4 bytes), because
# `Flux.withgradient` (train.jl:51-53) and FluxTraining (train_v2.jl:69) need a real-valued loss.  The rrules
# return the cotangent of the FIRST argument (the prediction), which is how both scripts use them; the cotangent of
# the target (the same kernels with the images swapped: both losses are symmetric) is computed when
# `ADMMTVLosses.TARGET_GRADIENT[] = true`.  `ssim` takes every window the reference takes (separable or not, per
# channel or shared) and `crop=false`.
module ADMMTVLosses

using CUDA, ChainRulesCore, Flux, LinearAlgebra

export gmsd, gmsd_loss, ssim, ssim_loss, ssim_loss_fast, ssim_kernel

const libadmmtv = get(ENV, "ADMMTV_LIB", joinpath(@__DIR__, "..", "libadmmtv.so"))
check(rc::Integer) = rc == 0 ? nothing :
    error(unsafe_string(ccall((:admmtv_strerror, libadmmtv), Cstring, (Cint,), rc)))
dev() = Cint(CUDA.deviceid(CUDA.device()))
strm() = CUDA.stream().handle

# Both losses are symmetric in their two images, so the pullback w.r.t. the SECOND argument is the same kernels with the
# images swapped.  Zygote evaluates every cotangent a rule returns, and both training scripts differentiate the
# prediction (first argument) only, so the second cotangent is computed on request:  ADMMTVLosses.TARGET_GRADIENT[] = true
const TARGET_GRADIENT = Ref(false)

# ---- GMSD (gmsd.jl:13-30) ----------------------------------------------------------------------
function gmsd_ws(x)
    M, N, C, B = size(x); n = Ref{Csize_t}(0)
    check(ccall((:admmtv_gmsd_workspace_bytes, libadmmtv), Cint, (Cint, Cint, Cint, Cint, Ref{Csize_t}), M, N, C, B, n))
    CUDA.zeros(UInt8, max(Int(n[]), 256))
end

function gmsd_fwd(x::CuArray{Float32,4}, y::CuArray{Float32,4}, t::Float32, α::Float32)
    M, N, C, B = size(x); ws = gmsd_ws(x); out = CUDA.zeros(Float32, 1)
    GC.@preserve x y out ws begin
        check(ccall((:admmtv_gmsd_forward, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cfloat, Cfloat, CuPtr{Cfloat}, CuPtr{Cvoid}, Ptr{Cvoid}),
                    M, N, C, B, dev(), pointer(x), pointer(y), t, α, pointer(out), pointer(ws), strm()))
    end
    out, ws
end

# gmsd.jl:13: `reduction` is applied to the (1,1,1,B) per-image deviations.  The device kernel returns their MEAN;
# `sum` is that times B.  Other reductions are not fused (the reference's scripts only use the default).
reduction_scale(::typeof(Flux.mean), B) = 1f0
reduction_scale(::typeof(sum), B) = Float32(B)
reduction_scale(f, B) = error("ADMMTVLosses.gmsd: reduction $(f) is not supported by the fused kernel (Flux.mean and sum are)")

function gmsd(x::CuArray{Float32,4}, y::CuArray{Float32,4}, t::Float32=0.0026f0, α::Float32=0.f0, reduction::Function=Flux.mean)
    reduction_scale(reduction, size(x, 4)) * only(Array(gmsd_fwd(x, y, t, α)[1]))
end
gmsd_loss(x::CuArray{Float32}, args...; kws...) = gmsd(x, args...; kws...)

function ChainRulesCore.rrule(::typeof(gmsd), x::CuArray{Float32,4}, y::CuArray{Float32,4}, t::Float32=0.0026f0, α::Float32=0.f0,
                              reduction::Function=Flux.mean)
    sc = reduction_scale(reduction, size(x, 4))
    out, ws = gmsd_fwd(x, y, t, α)
    function pullback(l̄)
        M, N, C, B = size(x); x̄ = similar(x); lb = CuArray(Float32[sc * unthunk(l̄)])
        GC.@preserve x y lb ws x̄ begin
            check(ccall((:admmtv_gmsd_backward, libadmmtv), Cint,
                        (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cfloat, Cfloat, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                        M, N, C, B, dev(), pointer(x), pointer(y), t, α, pointer(lb), pointer(ws), pointer(x̄), strm()))
        end
        ȳ = NoTangent()
        if TARGET_GRADIENT[]      # gmsd(x, y) = gmsd(y, x) (gmsd.jl:5-10): the swapped call
            _, ws2 = gmsd_fwd(y, x, t, α); ȳ = similar(y)
            GC.@preserve x y lb ws2 ȳ begin
                check(ccall((:admmtv_gmsd_backward, libadmmtv), Cint,
                            (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cfloat, Cfloat, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                            M, N, C, B, dev(), pointer(y), pointer(x), t, α, pointer(lb), pointer(ws2), pointer(ȳ), strm()))
            end
        end
        (NoTangent(), x̄, ȳ, NoTangent(), NoTangent(), NoTangent())
    end
    sc * only(Array(out)), pullback
end

# ---- SSIM (ssim.jl:84-164) ---------------------------------------------------------------------
# A window of ssim.jl:84 in the form the C ABI takes.  `taps`: equal separable taps (nothing = the 11-tap Gaussian of
# ssim.jl:6-17) -> the unrolled kernels admmtv_ssim_forward / _backward.  Otherwise `u` (L1 x R) and `v` (L2 x R) hold
# the R separable terms of an arbitrary (L1, L2) window, W = u * v' -> admmtv_ssim_window_forward / _backward
# (column-major (L, R) is exactly the row-major [R][L] layout of the header).
struct Window
    general::Bool
    taps::Union{Nothing,Vector{Float32}}
    u::Matrix{Float32}
    v::Matrix{Float32}
    L1::Int
    L2::Int
end
Window(taps::Nothing) = Window(false, nothing, zeros(Float32, 0, 0), zeros(Float32, 0, 0), 11, 11)
Window(taps::Vector{Float32}) = Window(false, taps, zeros(Float32, 0, 0), zeros(Float32, 0, 0), length(taps), length(taps))

function ssim_fwd(x, y, w::Window, peakval, as_loss::Bool, with_grad::Bool)
    M, N, C, B = size(x); n = Ref{Csize_t}(0); out = CUDA.zeros(Float32, 1)
    if w.general
        u = w.u; v = w.v; R = size(u, 2)
        check(ccall((:admmtv_ssim_window_workspace_bytes, libadmmtv), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ref{Csize_t}),
                    M, N, C, B, w.L1, w.L2, with_grad, n))
        ws = CUDA.zeros(UInt8, max(Int(n[]), 256))
        GC.@preserve u v x y out ws check(ccall((:admmtv_ssim_window_forward, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Cfloat, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                    M, N, C, B, dev(), pointer(x), pointer(y), pointer(u), pointer(v), w.L1, w.L2, R, Float32(peakval), as_loss, pointer(out), pointer(ws), with_grad, strm()))
        return out, ws
    end
    taps = w.taps
    L = taps === nothing ? 0 : length(taps); tp = taps === nothing ? C_NULL : pointer(taps)
    check(ccall((:admmtv_ssim_workspace_bytes, libadmmtv), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Ref{Csize_t}),
                M, N, C, B, L, with_grad, n))
    ws = CUDA.zeros(UInt8, max(Int(n[]), 256))
    GC.@preserve taps x y out ws check(ccall((:admmtv_ssim_forward, libadmmtv), Cint,
                (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Cint, Cfloat, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                M, N, C, B, dev(), pointer(x), pointer(y), tp, L, Float32(peakval), as_loss, pointer(out), pointer(ws), with_grad, strm()))
    out, ws
end

# cotangent of the FIRST image of a with_grad forward call on (x, y)
function ssim_bwd(x, y, w::Window, as_loss::Bool, ob, ws)
    M, N, C, B = size(x); x̄ = similar(x)
    if w.general
        u = w.u; v = w.v; R = size(u, 2)
        GC.@preserve u v x y ob ws x̄ check(ccall((:admmtv_ssim_window_backward, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                    M, N, C, B, dev(), pointer(x), pointer(y), pointer(u), pointer(v), w.L1, w.L2, R, as_loss, pointer(ob), pointer(ws), pointer(x̄), strm()))
        return x̄
    end
    taps = w.taps
    L = taps === nothing ? 0 : length(taps); tp = taps === nothing ? C_NULL : pointer(taps)
    GC.@preserve taps x y ob ws x̄ check(ccall((:admmtv_ssim_backward, libadmmtv), Cint,
                (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                M, N, C, B, dev(), pointer(x), pointer(y), tp, L, as_loss, pointer(ob), pointer(ws), pointer(x̄), strm()))
    x̄
end

_ssim(x, y, w::Window, peakval, as_loss) = only(Array(ssim_fwd(x, y, w, peakval, as_loss, false)[1]))
function ChainRulesCore.rrule(::typeof(_ssim), x, y, w::Window, peakval, as_loss)
    out, ws = ssim_fwd(x, y, w, peakval, as_loss, true)
    function pullback(ō)
        ob = CuArray(Float32[unthunk(ō)])
        x̄ = ssim_bwd(x, y, w, as_loss, ob, ws)
        ȳ = NoTangent()
        if TARGET_GRADIENT[]
            _, ws2 = ssim_fwd(y, x, w, peakval, as_loss, true)
            ȳ = ssim_bwd(y, x, w, as_loss, ob, ws2)
        end
        (NoTangent(), x̄, ȳ, NoTangent(), NoTangent(), NoTangent())
    end
    only(Array(out)), pullback
end

# NNlib.pad_symmetric(x, (lo1, hi1, lo2, hi2)) (ssim.jl:108-109, crop = false) and its pullback on the device
function pad_symmetric_dev(x::CuArray{Float32,4}, pads::NTuple{4,Int})
    M, N, C, B = size(x); lo1, hi1, lo2, hi2 = pads
    out = similar(x, M + lo1 + hi1, N + lo2 + hi2, C, B)
    GC.@preserve x out check(ccall((:admmtv_pad_symmetric, libadmmtv), Cint,
                (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cvoid}),
                M, N, C * B, lo1, hi1, lo2, hi2, dev(), pointer(x), pointer(out), strm()))
    out
end
function ChainRulesCore.rrule(::typeof(pad_symmetric_dev), x::CuArray{Float32,4}, pads::NTuple{4,Int})
    M, N, C, B = size(x); lo1, hi1, lo2, hi2 = pads
    function pullback(ō)
        ob = CuArray{Float32,4}(unthunk(ō)); x̄ = similar(x)
        GC.@preserve ob x̄ check(ccall((:admmtv_pad_symmetric_adjoint, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cvoid}),
                    M, N, C * B, lo1, hi1, lo2, hi2, dev(), pointer(ob), pointer(x̄), strm()))
        (NoTangent(), x̄, NoTangent())
    end
    pad_symmetric_dev(x, pads), pullback
end

# ssim.jl:25-47: the 11 x 11 Gaussian (σ = 1.5) as an (11,11,1,1) array -- kept so that call sites which build the window
# themselves (`ssim(x, y, ssim_kernel(x))`, ssim.jl:84) keep working; `windows` recognises it as separable.
const SSIM_TAPS = let g = [exp(-Float32(i)^2 / (2 * 1.5f0^2)) for i in -5:5]
    g ./ sum(g)
end
ssim_kernel(::Type{T}, N::Integer) where {T} = N == 4 ? reshape(T.(SSIM_TAPS * SSIM_TAPS'), 11, 11, 1, 1) :
    throw("ADMMTVLosses.ssim is implemented for 4D inputs, dimension=$N provided.")
ssim_kernel(x::AbstractArray{T,N}) where {T,N} = ssim_kernel(T, N)
ChainRulesCore.@non_differentiable ssim_kernel(T::Any, N::Any)
ChainRulesCore.@non_differentiable ssim_kernel(x::Any)

# One (L1, L2) window -> Window: its singular value decomposition gives the separable terms; a rank-one window with equal
# factors (the Gaussian, the box of ssim_loss_fast) takes the unrolled tap kernels.
function window_of(K1::AbstractMatrix)
    W = Float64.(Array(K1)); L1, L2 = size(W)
    (1 <= L1 <= 11 && 1 <= L2 <= 11) || error("ADMMTVLosses.ssim: the window must be at most 11 x 11")
    F = svd(W)
    keep = [r for r in 1:length(F.S) if F.S[r] > 1e-7 * F.S[1]]
    isempty(keep) && (keep = [1])
    if length(keep) == 1 && L1 == L2
        su = F.U[:, 1] .* sqrt(F.S[1]); sv = F.V[:, 1] .* sqrt(F.S[1])
        if sum(su) < 0
            su = -su; sv = -sv
        end
        if maximum(abs.(su .- sv)) <= 1e-7 * maximum(abs.(su))
            return Window(Float32.(su))
        end
    end
    u = Float32.(F.U[:, keep] .* F.S[keep]')
    v = Float32.(F.V[:, keep])
    Window(true, nothing, Matrix(u), Matrix(v), L1, L2)
end

# kernel_ref (ssim.jl:84) -> Vector of (channel range, Window): `nothing` (default window), a vector of 1-D taps, a 2-D
# window, or the reference's 4-D (L1, L2, 1, 1|C) array; distinct per-channel windows are evaluated channel by channel.
windows(::Nothing, C) = [(1:C, Window(nothing))]
windows(t::AbstractVector, C) = [(1:C, Window(Float32.(Array(t))))]
windows(k::AbstractMatrix, C) = [(1:C, window_of(k))]
function windows(k::AbstractArray{<:Any,4}, C)
    K = Array(k)
    (size(K, 3) == 1 && (size(K, 4) == 1 || size(K, 4) == C)) || error("ADMMTVLosses.ssim: the window must be (L1,L2,1,1) or (L1,L2,1,$C)")
    if size(K, 4) == 1 || all(c -> K[:, :, 1, c] == K[:, :, 1, 1], 1:size(K, 4))
        return [(1:C, window_of(K[:, :, 1, 1]))]
    end
    [(c:c, window_of(K[:, :, 1, c])) for c in 1:C]
end
ChainRulesCore.@non_differentiable windows(k::Any, C::Any)

# ssim.jl:104-110: calc_padding of Flux's conv.jl, (cld, fld) of L - 1 per dimension
same_pads(w::Window) = (cld(w.L1 - 1, 2), fld(w.L1 - 1, 2), cld(w.L2 - 1, 2), fld(w.L2 - 1, 2))
ChainRulesCore.@non_differentiable same_pads(w::Any)

function ssim_impl(x::CuArray{Float32,4}, y::CuArray{Float32,4}, kernel_ref, peakval, crop::Bool, as_loss::Bool)
    size(x) == size(y) || throw(DimensionMismatch("loss function expects size(ŷ) = $(size(y)) but is size $(size(x))"))   # ssim.jl:56-62
    parts = windows(kernel_ref, size(x, 3))
    vals = map(parts) do (cs, w)
        xs = length(parts) == 1 ? x : x[:, :, cs, :]
        ys = length(parts) == 1 ? y : y[:, :, cs, :]
        if !crop
            xs = pad_symmetric_dev(xs, same_pads(w)); ys = pad_symmetric_dev(ys, same_pads(w))
        end
        _ssim(xs, ys, w, peakval, as_loss)
    end
    # equal-sized maps: the mean over (1,2,3) then the batch is the mean of the per-channel means (ssim.jl:122-123)
    sum(vals) / length(vals)
end

# `dims` is accepted and ignored, exactly as in the reference (ssim.jl:84-124 never reads it).
ssim(x::CuArray{Float32,4}, y::CuArray{Float32,4}, kernel_ref=nothing; peakval=1f0, crop=true, dims=:) =
    ssim_impl(x, y, kernel_ref, peakval, crop, false)
ssim_loss(x::CuArray{Float32,4}, y::CuArray{Float32,4}, kernel_ref=nothing; peakval=1f0, crop=true, dims=:) =
    ssim_impl(x, y, kernel_ref, peakval, crop, true)
ssim_loss_fast(x::CuArray{Float32,4}, y::CuArray{Float32,4}; kernel_length=5, kws...) =
    ssim_loss(x, y, fill(1f0 / kernel_length, kernel_length); kws...)

end # module
