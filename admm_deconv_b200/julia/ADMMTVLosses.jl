# ADMMTVLosses.jl -- Julia drop-in for /root/reference/src/metrics/gmsd.jl (+ iqa_utils.jl) and
# src/metrics/ssim.jl over the C ABI of include/admmtv_loss.h (libadmmtv.so, sm_100a kernels).
#
# STATUS: written to the ABI, NOT EXECUTED (no Julia toolchain in the build image or on the GPU box); the
# identical ABI is exercised from Python (admm_deconv_b200/losses.py, tests/test_gpu_losses.py).  Usage:
#
#     # src/train.jl, src/train_v2.jl:
#     # replace  include("metrics/gmsd.jl"); include("metrics/ssim.jl")  by  include("ADMMTVLosses.jl"); using .ADMMTVLosses
#
# Same names and argument order: gmsd(x, y, t, α), gmsd_loss, ssim(x, y; peakval), ssim_loss, ssim_loss_fast.
# The reduction happens on the device; the value is read back as a Float32 scalar (4 bytes), because
# `Flux.withgradient` (train.jl:51-53) and FluxTraining (train_v2.jl:69) need a real-valued loss.  The rrules
# return the cotangent of the FIRST argument (the prediction) and NoTangent for the target, which is how both
# scripts use them.
module ADMMTVLosses

using CUDA, ChainRulesCore

export gmsd, gmsd_loss, ssim, ssim_loss, ssim_loss_fast

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
    check(ccall((:admmtv_gmsd_forward, libadmmtv), Cint,
                (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cfloat, Cfloat, CuPtr{Cfloat}, CuPtr{Cvoid}, Ptr{Cvoid}),
                M, N, C, B, dev(), x, y, t, α, out, ws, strm()))
    out, ws
end

gmsd(x::CuArray{Float32,4}, y::CuArray{Float32,4}, t::Float32=0.0026f0, α::Float32=0.f0) = only(Array(gmsd_fwd(x, y, t, α)[1]))
gmsd_loss(x, args...; kws...) = gmsd(x, args...; kws...)

function ChainRulesCore.rrule(::typeof(gmsd), x::CuArray{Float32,4}, y::CuArray{Float32,4}, t::Float32=0.0026f0, α::Float32=0.f0)
    out, ws = gmsd_fwd(x, y, t, α)
    function pullback(l̄)
        M, N, C, B = size(x); x̄ = similar(x); lb = CuArray(Float32[unthunk(l̄)])
        check(ccall((:admmtv_gmsd_backward, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cfloat, Cfloat, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                    M, N, C, B, dev(), x, y, t, α, lb, ws, x̄, strm()))
        (NoTangent(), x̄, NoTangent(), NoTangent(), NoTangent())
    end
    only(Array(out)), pullback
end

# ---- SSIM (ssim.jl:84-164) ---------------------------------------------------------------------
# `taps`: the 1-D taps of a separable window (nothing = the 11-tap Gaussian of ssim.jl:6-17)
function ssim_fwd(x, y, taps, peakval, as_loss::Bool, with_grad::Bool)
    M, N, C, B = size(x); n = Ref{Csize_t}(0)
    L = taps === nothing ? 0 : length(taps); tp = taps === nothing ? C_NULL : pointer(taps)
    check(ccall((:admmtv_ssim_workspace_bytes, libadmmtv), Cint, (Cint, Cint, Cint, Cint, Cint, Cint, Ref{Csize_t}),
                M, N, C, B, L, with_grad, n))
    ws = CUDA.zeros(UInt8, max(Int(n[]), 256)); out = CUDA.zeros(Float32, 1)
    GC.@preserve taps check(ccall((:admmtv_ssim_forward, libadmmtv), Cint,
                (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Cint, Cfloat, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                M, N, C, B, dev(), x, y, tp, L, Float32(peakval), as_loss, out, ws, with_grad, strm()))
    out, ws
end

_ssim(x, y, taps, peakval, as_loss) = only(Array(ssim_fwd(x, y, taps, peakval, as_loss, false)[1]))
function ChainRulesCore.rrule(::typeof(_ssim), x, y, taps, peakval, as_loss)
    out, ws = ssim_fwd(x, y, taps, peakval, as_loss, true)
    function pullback(ō)
        M, N, C, B = size(x); x̄ = similar(x); ob = CuArray(Float32[unthunk(ō)])
        L = taps === nothing ? 0 : length(taps); tp = taps === nothing ? C_NULL : pointer(taps)
        GC.@preserve taps check(ccall((:admmtv_ssim_backward, libadmmtv), Cint,
                    (Cint, Cint, Cint, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Ptr{Cfloat}, Cint, Cint, CuPtr{Cfloat}, CuPtr{Cvoid}, CuPtr{Cfloat}, Ptr{Cvoid}),
                    M, N, C, B, dev(), x, y, tp, L, as_loss, ob, ws, x̄, strm()))
        (NoTangent(), x̄, NoTangent(), NoTangent(), NoTangent(), NoTangent())
    end
    only(Array(out)), pullback
end

ssim(x::CuArray{Float32,4}, y::CuArray{Float32,4}, taps=nothing; peakval=1f0) = _ssim(x, y, taps, peakval, false)
ssim_loss(x::CuArray{Float32,4}, y::CuArray{Float32,4}, taps=nothing; peakval=1f0) = _ssim(x, y, taps, peakval, true)
ssim_loss_fast(x, y; kernel_length=5, kws...) = ssim_loss(x, y, fill(1f0 / kernel_length, kernel_length); kws...)

end # module
