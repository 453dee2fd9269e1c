# ADMMTV.jl -- Julia drop-in for /root/reference/src/layers/deconv_admm.jl + src/ops/ops.jl over
# the C ABI of include/admmtv.h (libadmmtv.so, hand-written sm_100a kernels).
#
# STATUS: written to the ABI, NOT EXECUTED -- there is no Julia toolchain in the build image nor on
# the GPU box (SURVEY.md 8b "Harness reality").  The identical ABI is exercised by the Python ctypes
# harness (admm_deconv_b200/_lib.py, tests/).  Usage in the reference:
#
#     # src/nets/net_build.jl, src/train.jl, src/train_v2.jl, src/ADMM_Deconv.jl:
#     # replace   include("../layers/deconv_admm.jl")   by   include("ADMMTV.jl"); using .ADMMTV
#
# Same struct names, field names and order (σ, weight, bias, λ, ρ, iters, iso, creg), both
# constructor forms per variant, the same `Flux.@layer ... trainable=` tuples, the `Admm` union,
# the call `(d::Admm)(x)` and the operator `tvd_fft(y, λ, ρ, h, isotropic, maxit)`.
# No CUDA.jl broadcast kernel runs on the path: clamp, bias and σ are fused into the shim calls.
module ADMMTV

using Flux, CUDA, ChainRulesCore

export ADMMDeconv, ADMMDeconvF1, ADMMDeconvF2, ADMMDeconvF3, Admm, tvd_fft, tvd_fft_gpu, tvd_fft_cpu, nograd_repeat!

const libadmmtv = get(ENV, "ADMMTV_LIB", joinpath(@__DIR__, "..", "libadmmtv.so"))

# struct admmtv_desc (include/admmtv.h) -- 14 x 4 bytes
struct Desc
    M::Int32; N::Int32; P::Int32; B::Int32
    kh::Int32; kw::Int32
    iters::Int32; iso::Int32; activation::Int32; has_bias::Int32
    device::Int32; flags::Int32
    creg::Float32; groups::Int32
end

const FLAG_NO_CLAMP = Int32(1)
const FLAG_NOGRAD_REPEAT = Int32(2)
const FLAG_PER_ITER_PARAMS = Int32(64)   # EXTENSION: λ, ρ hold `iters` values, entry k for iteration k (include/admmtv.h)
per_iter_flag(λ, iters) = length(λ) > 1 ? (length(λ) == iters ? FLAG_PER_ITER_PARAMS :
    error("ADMMTV: per-iteration parameters need exactly `iters` values")) : Int32(0)

# train.jl:10 declares `Zygote.@nograd CUDA.repeat` (= Base.repeat): under it Zygote drops the gradient through
# `h = repeat(h,1,1,1,B)` (ops.jl:153), i.e. ∂weight loses its spatial H^T y path.  train_v2.jl / ADMM_Deconv.jl do not.
# That is a process-global switch in the reference, so it is one here: call `ADMMTV.nograd_repeat!(true)` where the
# reference script has that line (it sets ADMMTV_FLAG_NOGRAD_REPEAT on every subsequent layer call and pullback).
const NOGRAD_REPEAT = Ref(false)
nograd_repeat!(on::Bool=true) = (NOGRAD_REPEAT[] = on)
grad_flags() = NOGRAD_REPEAT[] ? FLAG_NOGRAD_REPEAT : Int32(0)

# σ -> ADMMTV_ACT_* (the activations net_build.jl uses: identity, relu, relu6, relu1 (:8))
relu1(x) = min(max(0, x), 1)
act_code(::typeof(identity)) = Int32(0)
act_code(::typeof(Flux.relu)) = Int32(1)
act_code(::typeof(Flux.relu6)) = Int32(2)
act_code(f) = nameof(f) === :relu1 ? Int32(3) :
    error("ADMMTV: activation $(f) is not fused in the shim (identity, relu, relu6, relu1 are)")

check(rc::Integer) = rc == 0 ? nothing :
    error(unsafe_string(ccall((:admmtv_strerror, libadmmtv), Cstring, (Cint,), rc)))

function make_desc(y, h, iters, iso, act, has_bias, flags, creg)
    M, N, P, B = size(y)
    kh, kw = isempty(h) ? (0, 0) : (size(h, 1), size(h, 2))
    Desc(M, N, P, B, kh, kw, iters, iso ? 1 : 0, act, has_bias ? 1 : 0, CUDA.deviceid(CUDA.device()), flags,
         Float32(creg), 0)
end

function workspace_bytes(d::Desc)
    f = Ref{Csize_t}(0); c = Ref{Csize_t}(0); b = Ref{Csize_t}(0)
    check(ccall((:admmtv_workspace_bytes, libadmmtv), Cint, (Ref{Desc}, Ref{Csize_t}, Ref{Csize_t}, Ref{Csize_t}), d, f, c, b))
    Int(f[]), Int(c[]), Int(b[])
end

ptr_or_null(a::CuArray{Float32}) = isempty(a) ? CU_NULL : pointer(a)
ptr_or_null(::Any) = CU_NULL            # `false` bias, empty weight

# ---- raw calls -------------------------------------------------------------------------------------
# λ, ρ, h are clamped IN PLACE by the shim: that is the reference's write-back at
# deconv_admm.jl:216-219 (the Julia arrays passed in are the layer's own fields).
# The ccalls only ENQUEUE on the task's stream: every array whose pointer crosses the boundary is kept rooted with
# GC.@preserve for the duration of the call, and the caller-owned workspace is returned to the caller, which holds it
# until the stream work has completed (CUDA.jl's pool frees stream-ordered, so dropping it after the call that enqueued
# the work on the same stream is safe; GC finalisation in the middle of the ccall is not).
function forward!(d::Desc, y, h, λ, ρ, bias; ckpt::Bool)
    fwd, ck, _ = workspace_bytes(d)
    ws = CuArray{UInt8}(undef, fwd)                       # caller-owned: CUDA.jl's pool accounts for it
    ckb = ckpt ? CuArray{UInt8}(undef, ck) : nothing
    x = similar(y)
    GC.@preserve y h λ ρ bias x ws ckb begin
        check(ccall((:admmtv_forward, libadmmtv), Cint,
            (Ref{Desc}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat},
             CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
            d, pointer(y), ptr_or_null(h), pointer(λ), pointer(ρ), ptr_or_null(bias), pointer(x),
            pointer(ws), ckb === nothing ? CU_NULL : pointer(ckb), CUDA.stream().handle))
    end
    x, ckb, ws
end

function backward!(d::Desc, x̄, x, y, h, λ, ρ, ckb)
    _, _, bwd = workspace_bytes(d)
    ws = CuArray{UInt8}(undef, bwd)
    ȳ = similar(y); h̄ = similar(h); λ̄ = similar(λ); ρ̄ = similar(ρ)
    b̄ = d.has_bias == 1 ? CUDA.zeros(Float32, 1) : nothing
    GC.@preserve x̄ x y h λ ρ ckb ȳ h̄ λ̄ ρ̄ b̄ ws begin
        check(ccall((:admmtv_backward, libadmmtv), Cint,
            (Ref{Desc}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cvoid},
             CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cvoid}, Ptr{Cvoid}),
            d, pointer(x̄), pointer(x), pointer(y), ptr_or_null(h), pointer(λ), pointer(ρ), pointer(ckb),
            pointer(ȳ), ptr_or_null(h̄), pointer(λ̄), pointer(ρ̄), b̄ === nothing ? CU_NULL : pointer(b̄),
            pointer(ws), CUDA.stream().handle))
    end
    ȳ, h̄, λ̄, ρ̄, b̄, ws
end

# ---- the operator (ops.jl:181-188) -----------------------------------------------------------------
# Differentiable core shared by tvd_fft and the layer call.
function admm_call(y::CuArray{Float32,4}, λ::CuArray{Float32,1}, ρ::CuArray{Float32,1}, h, bias,
                   iso::Bool, iters::Integer, act::Int32, creg, flags::Int32)
    d = make_desc(y, h, iters, iso, act, bias isa CuArray, flags | per_iter_flag(λ, iters), creg)
    x, _, ws = forward!(d, y, h, λ, ρ, bias; ckpt=false)
    CUDA.unsafe_free!(ws)          # stream-ordered: returns to the pool after the work enqueued above
    x
end

function ChainRulesCore.rrule(::typeof(admm_call), y, λ, ρ, h, bias, iso, iters, act, creg, flags)
    d = make_desc(y, h, iters, iso, act, bias isa CuArray, flags | grad_flags() | per_iter_flag(λ, iters), creg)
    x, ckb, ws = forward!(d, y, h, λ, ρ, bias; ckpt=true)
    CUDA.unsafe_free!(ws)
    function admm_pullback(x̄)
        ȳ, h̄, λ̄, ρ̄, b̄, wsb = backward!(d, CuArray{Float32,4}(unthunk(x̄)), x, y, h, λ, ρ, ckb)
        CUDA.unsafe_free!(wsb)
        (NoTangent(), ȳ, λ̄, ρ̄, isempty(h) ? NoTangent() : h̄, b̄ === nothing ? NoTangent() : b̄,
         NoTangent(), NoTangent(), NoTangent(), NoTangent(), NoTangent())
    end
    x, admm_pullback
end

function tvd_fft(y::CuArray{Float32,4}, λ::CuArray{Float32,1}, ρ::CuArray{Float32,1}=CuArray(Float32[1]),
                 h::CuArray{Float32}=CuArray{Float32}(undef, 0), isotropic=false, maxit=100)
    admm_call(y, λ, ρ, h, false, Bool(isotropic), maxit, Int32(0), 0f0, FLAG_NO_CLAMP)
end
const tvd_fft_gpu = tvd_fft     # tests/admm_deconv_test.jl:76 calls tvd_fft_gpu directly

# CPU `Array` arguments: the reference dispatches them to tvd_fft_cpu (ops.jl:183-187).  Here they take the host-buffer
# entry point admmtv_forward_host (copies inside; the arithmetic still runs on the GPU -- there is no CPU path).
# Not differentiable (the reference's CPU twin is only reachable from inference scripts).
function tvd_fft(y::Array{Float32,4}, λ::Array{Float32,1}, ρ::Array{Float32,1}=Float32[1],
                 h::Array{Float32}=Array{Float32}(undef, 0), isotropic=false, maxit=100)
    M, N, P, B = size(y)
    kh, kw = isempty(h) ? (0, 0) : (size(h, 1), size(h, 2))
    d = Desc(M, N, P, B, kh, kw, maxit, Bool(isotropic) ? 1 : 0, 0, 0, CUDA.deviceid(CUDA.device()), FLAG_NO_CLAMP, 0f0, 0)
    x = similar(y)
    λc, ρc, hc = copy(λ), copy(ρ), copy(h)       # IN/OUT in the C ABI (persisted clamp); NO_CLAMP leaves them untouched
    GC.@preserve y λc ρc hc x begin
        check(ccall((:admmtv_forward_host, libadmmtv), Cint,
            (Ref{Desc}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}),
            d, pointer(y), isempty(hc) ? C_NULL : pointer(hc), pointer(λc), pointer(ρc), C_NULL, pointer(x)))
    end
    x
end
const tvd_fft_cpu = tvd_fft     # ops.jl:17 (same call; runs on the GPU through the host-buffer entry point)

# ---- the layers (deconv_admm.jl:6-212) --------------------------------------------------------------
# Written out literally (the reference does the same at deconv_admm.jl:6-15,55 / 58-67,107 / 110-119,161 / 164-173,209):
# Flux.@layer checks `Meta.isexpr(trainable, :tuple)`, so the field tuple must appear as source text, not as an
# interpolated value.
mutable struct ADMMDeconvF1{F,A,N,V,M,B,C,D}
    σ::F
    weight::A
    bias::V
    λ::N
    ρ::M
    iters::B
    iso::C
    creg::D
end
# 8-positional form (w, σ, b, λ, ρ, iters, iso, creg)
ADMMDeconvF1(w::AbstractArray{T_,N_}, σ, b, lambda, rho, iters, iso, creg) where {T_,N_} =
    ADMMDeconvF1(σ, w, b, lambda, rho, iters, iso, creg)
Flux.@layer ADMMDeconvF1 trainable=(weight, bias, ρ,)

mutable struct ADMMDeconvF2{F,A,N,V,M,B,C,D}
    σ::F
    weight::A
    bias::V
    λ::N
    ρ::M
    iters::B
    iso::C
    creg::D
end
# 8-positional form (w, σ, b, λ, ρ, iters, iso, creg)
ADMMDeconvF2(w::AbstractArray{T_,N_}, σ, b, lambda, rho, iters, iso, creg) where {T_,N_} =
    ADMMDeconvF2(σ, w, b, lambda, rho, iters, iso, creg)
Flux.@layer ADMMDeconvF2 trainable=(weight, bias, λ,)

mutable struct ADMMDeconvF3{F,A,N,V,M,B,C,D}
    σ::F
    weight::A
    bias::V
    λ::N
    ρ::M
    iters::B
    iso::C
    creg::D
end
# 8-positional form (w, σ, b, λ, ρ, iters, iso, creg)
ADMMDeconvF3(w::AbstractArray{T_,N_}, σ, b, lambda, rho, iters, iso, creg) where {T_,N_} =
    ADMMDeconvF3(σ, w, b, lambda, rho, iters, iso, creg)
Flux.@layer ADMMDeconvF3 trainable=(weight, bias,)

mutable struct ADMMDeconv{F,A,N,V,M,B,C,D}
    σ::F
    weight::A
    bias::V
    λ::N
    ρ::M
    iters::B
    iso::C
    creg::D
end
# 8-positional form (w, σ, b, λ, ρ, iters, iso, creg)
ADMMDeconv(w::AbstractArray{T_,N_}, σ, b, lambda, rho, iters, iso, creg) where {T_,N_} =
    ADMMDeconv(σ, w, b, lambda, rho, iters, iso, creg)
Flux.@layer ADMMDeconv trainable=(weight, bias, λ, ρ,)

_weight(k, init, groups) = isempty(k) ? empty(ones(1)) : Flux.convfilter(k, 1 => 1; init=init, groups=groups)

function ADMMDeconv(k::NTuple{N,Integer}, num_it::Integer, σ=Flux.identity; iso::Bool=false,
                    init=Flux.glorot_uniform, groups=1, bias=false, creg::Number=0f0) where {N}
    weight = _weight(k, init, groups)
    λ = abs.(Flux.glorot_uniform(1)); ρ = abs.(Flux.glorot_uniform(1))
    ADMMDeconv(weight, σ, Flux.create_bias(weight, bias, 1), λ, ρ, num_it, iso, creg)
end
function ADMMDeconvF1(k::NTuple{N,Integer}, num_it::Integer, λ, σ=Flux.identity; iso::Bool=false,
                      init=Flux.glorot_uniform, groups=1, bias=false, creg::Number=0f0) where {N}
    @assert λ > 0f0 "Parameter λ must be greater than 0"
    weight = _weight(k, init, groups)
    ADMMDeconvF1(weight, σ, Flux.create_bias(weight, bias, 1), zeros(1) .+ λ, abs.(Flux.glorot_uniform(1)), num_it, iso, creg)
end
function ADMMDeconvF2(k::NTuple{N,Integer}, num_it::Integer, ρ, σ=Flux.identity; iso::Bool=false,
                      init=Flux.glorot_uniform, groups=1, bias=false, creg::Number=0f0) where {N}
    @assert ρ > 0 "Parameter ρ must be greater than 0"
    weight = _weight(k, init, groups)
    ADMMDeconvF2(weight, σ, Flux.create_bias(weight, bias, 1), abs.(Flux.glorot_uniform(1)), zeros(1) .+ ρ, num_it, iso, creg)
end
function ADMMDeconvF3(k::NTuple{N,Integer}, num_it::Integer, λ, ρ, σ=Flux.identity; iso::Bool=false,
                      init=Flux.glorot_uniform, groups=1, bias=false, creg::Number=0f0) where {N}
    @assert λ > 0 "Parameter λ must be greater than 0"
    @assert ρ > 0 "Parameter ρ must be greater than 0"
    weight = _weight(k, init, groups)
    ADMMDeconvF3(weight, σ, Flux.create_bias(weight, bias, 1), zeros(1) .+ λ, zeros(1) .+ ρ, num_it, iso, creg)
end

const Admm = Union{ADMMDeconv,ADMMDeconvF1,ADMMDeconvF2,ADMMDeconvF3}

# (d::Admm)(x), deconv_admm.jl:215-225.  The shim clamps d.λ, d.ρ, d.weight in place (the fields ARE
# the arrays passed), adds the bias and applies σ; after `gpu(model)` the fields are Float32 CuArrays.
function (d::Admm)(x::CuArray{Float32,4})
    admm_call(x, d.λ, d.ρ, d.weight, d.bias, d.iso, d.iters, act_code(d.σ), d.creg, Int32(0))
end
# CPU model (before `gpu(model)`): same semantics through the host-buffer entry point, parameters clamped in place
function (d::Admm)(x::Array{Float32,4})
    M, N, P, B = size(x)
    w = Array{Float32}(d.weight); λ = Array{Float32}(d.λ); ρ = Array{Float32}(d.ρ)
    kh, kw = isempty(w) ? (0, 0) : (size(w, 1), size(w, 2))
    hasb = d.bias isa AbstractArray
    bias = hasb ? Array{Float32}(d.bias) : Float32[]
    desc = Desc(M, N, P, B, kh, kw, d.iters, d.iso ? 1 : 0, act_code(d.σ), hasb ? 1 : 0, CUDA.deviceid(CUDA.device()),
                Int32(0), Float32(d.creg), 0)
    out = similar(x)
    GC.@preserve x w λ ρ bias out begin
        check(ccall((:admmtv_forward_host, libadmmtv), Cint,
            (Ref{Desc}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}),
            desc, pointer(x), isempty(w) ? C_NULL : pointer(w), pointer(λ), pointer(ρ), hasb ? pointer(bias) : C_NULL, pointer(out)))
    end
    d.λ = oftype(d.λ, λ); d.ρ = oftype(d.ρ, ρ)                 # deconv_admm.jl:216-219: the clamp is persisted
    isempty(w) || (d.weight = oftype(d.weight, w))
    out
end

end # module
