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
# return the cotangent of the FIRST argument (the prediction) and NoTangent for the target, which is how both
# scripts use them.
module ADMMTVLosses

using CUDA, ChainRulesCore, Flux

export gmsd, gmsd_loss, ssim, ssim_loss, ssim_loss_fast, ssim_kernel

const libadmmtv = get(ENV, "ADMMTV_LIB", joinpath(@__DIR__, "..", "libadmmtv.so"))
check(rc::Integer) = rc == 0 ? nothing :
    error(unsafe_string(ccall((:admmtv_strerror, libadmmtv), Cstring, (Cint,), rc)))
dev() = Cint(CUDA.deviceid(CUDA.device()))
strm() = CUDA.stream().handle

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
        (NoTangent(), x̄, NoTangent(), NoTangent(), NoTangent(), NoTangent())
    end
    sc * only(Array(out)), pullback
end

# ---- SSIM (ssim.jl:84-164) ---------------------------------------------------------------------
# `taps`: the 1-D taps of a separable window (nothing = the 11-tap Gaussian of ssim.jl:6-17)
function ssim_fwd(x, y, taps, peakval, as_loss::Bool, with_grad::Bool)
    M, N, C, B = size(x); n = Ref{Csize_t}(0)
    L = taps === nothing ? 0 : length(taps); tp = taps === nothing ? C_NULL : pointer(taps)
    check(ccall((:admmtv_ssim_workspace_bytes, libadmmtv), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Ref{Csize_t}),
                M, N, C, B, L, with_grad, n))
    ws = CUDA.zeros(UInt8, max(Int(n[]), 256)); out = CUDA.zeros(Float32, 1)
    GC.@preserve taps x y out ws check(ccall((:admmtv_ssim_forward, libadmmtv), Cint,
                (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Cint, Cfloat, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                M, N, C, B, dev(), pointer(x), pointer(y), tp, L, Float32(peakval), as_loss, pointer(out), pointer(ws), with_grad, strm()))
    out, ws
end

_ssim(x, y, taps, peakval, as_loss) = only(Array(ssim_fwd(x, y, taps, peakval, as_loss, false)[1]))
function ChainRulesCore.rrule(::typeof(_ssim), x, y, taps, peakval, as_loss)
    out, ws = ssim_fwd(x, y, taps, peakval, as_loss, true)
    function pullback(ō)
        M, N, C, B = size(x); x̄ = similar(x); ob = CuArray(Float32[unthunk(ō)])
        L = taps === nothing ? 0 : length(taps); tp = taps === nothing ? C_NULL : pointer(taps)
        GC.@preserve taps x y ob ws x̄ check(ccall((:admmtv_ssim_backward, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                    M, N, C, B, dev(), pointer(x), pointer(y), tp, L, as_loss, pointer(ob), pointer(ws), pointer(x̄), strm()))
        (NoTangent(), x̄, NoTangent(), NoTangent(), NoTangent(), NoTangent())
    end
    only(Array(out)), pullback
end

# ssim.jl:25-47: the 11 x 11 Gaussian (σ = 1.5) as an (11,11,1,1) array -- kept so that call sites which build the window
# themselves (`ssim(x, y, ssim_kernel(x))`, ssim.jl:84) keep working; window_taps recognises it and every other
# separable window.
const SSIM_TAPS = let g = [exp(-Float32(i)^2 / (2 * 1.5f0^2)) for i in -5:5]
    g ./ sum(g)
end
ssim_kernel(::Type{T}, N::Integer) where {T} = N == 4 ? reshape(T.(SSIM_TAPS * SSIM_TAPS'), 11, 11, 1, 1) :
    throw("ADMMTVLosses.ssim is implemented for 4D inputs, dimension=$N provided.")
ssim_kernel(x::AbstractArray{T,N}) where {T,N} = ssim_kernel(T, N)
ChainRulesCore.@non_differentiable ssim_kernel(T::Any, N::Any)
ChainRulesCore.@non_differentiable ssim_kernel(x::Any)

# kernel_ref (ssim.jl:84): `nothing` (default window), a vector of 1-D taps, or the reference's 4-D (L,L,1,1|C) window.
# The fused kernels run SEPARABLE windows that are the same for every channel: a 4-D window is reduced to its taps
# t (K = t tᵀ, sum(t) = 1 ⇒ t = row sums) and checked; anything else is rejected rather than silently approximated.
window_taps(::Nothing) = nothing
window_taps(t::AbstractVector) = Float32.(Array(t))
function window_taps(k::AbstractArray{<:Any,4})
    K = Float32.(Array(k))
    size(K, 1) == size(K, 2) && size(K, 3) == 1 || error("ADMMTVLosses.ssim: the window must be (L,L,1,C)")
    K1 = K[:, :, 1, 1]
    all(c -> K[:, :, 1, c] ≈ K1, 1:size(K, 4)) || error("ADMMTVLosses.ssim: per-channel windows are not supported")
    s = sum(K1); t = vec(sum(K1, dims=2)) ./ sqrt(s)
    isapprox(t * t', K1; rtol=1f-4) || error("ADMMTVLosses.ssim: non-separable windows are not supported")
    t
end
ChainRulesCore.@non_differentiable window_taps(k::Any)
function check_ssim_kws(crop, dims)
    crop === true || error("ADMMTVLosses.ssim: crop=false is not supported")
    dims === Colon() || error("ADMMTVLosses.ssim: only dims=: is supported")
end

function ssim(x::CuArray{Float32,4}, y::CuArray{Float32,4}, kernel_ref=nothing; peakval=1f0, crop=true, dims=:)
    check_ssim_kws(crop, dims)
    _ssim(x, y, window_taps(kernel_ref), peakval, false)
end
function ssim_loss(x::CuArray{Float32,4}, y::CuArray{Float32,4}, kernel_ref=nothing; peakval=1f0, crop=true, dims=:)
    check_ssim_kws(crop, dims)
    _ssim(x, y, window_taps(kernel_ref), peakval, true)
end
ssim_loss_fast(x::CuArray{Float32,4}, y::CuArray{Float32,4}; kernel_length=5, kws...) =
    ssim_loss(x, y, fill(1f0 / kernel_length, kernel_length); kws...)

end # module
