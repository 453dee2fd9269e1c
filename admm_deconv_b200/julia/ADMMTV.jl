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

export ADMMDeconv, ADMMDeconvF1, ADMMDeconvF2, ADMMDeconvF3, Admm, tvd_fft, tvd_fft_gpu

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
function forward!(d::Desc, y, h, λ, ρ, bias; ckpt::Bool)
    fwd, ck, _ = workspace_bytes(d)
    ws = CuArray{UInt8}(undef, fwd)                       # caller-owned: CUDA.jl's pool accounts for it
    ckb = ckpt ? CuArray{UInt8}(undef, ck) : nothing
    x = similar(y)
    check(ccall((:admmtv_forward, libadmmtv), Cint,
        (Ref{Desc}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat},
         CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
        d, pointer(y), ptr_or_null(h), pointer(λ), pointer(ρ), ptr_or_null(bias), pointer(x),
        pointer(ws), ckb === nothing ? CU_NULL : pointer(ckb), CUDA.stream().handle))
    x, ckb
end

function backward!(d::Desc, x̄, x, y, h, λ, ρ, ckb)
    _, _, bwd = workspace_bytes(d)
    ws = CuArray{UInt8}(undef, bwd)
    ȳ = similar(y); h̄ = similar(h); λ̄ = similar(λ); ρ̄ = similar(ρ)
    b̄ = d.has_bias == 1 ? CUDA.zeros(Float32, 1) : nothing
    check(ccall((:admmtv_backward, libadmmtv), Cint,
        (Ref{Desc}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cvoid},
         CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cfloat}, CuPtr{Cvoid}, Ptr{Cvoid}),
        d, pointer(x̄), pointer(x), pointer(y), ptr_or_null(h), pointer(λ), pointer(ρ), pointer(ckb),
        pointer(ȳ), ptr_or_null(h̄), pointer(λ̄), pointer(ρ̄), b̄ === nothing ? CU_NULL : pointer(b̄),
        pointer(ws), CUDA.stream().handle))
    ȳ, h̄, λ̄, ρ̄, b̄
end

# ---- the operator (ops.jl:181-188) -----------------------------------------------------------------
# Differentiable core shared by tvd_fft and the layer call.
function admm_call(y::CuArray{Float32,4}, λ::CuArray{Float32,1}, ρ::CuArray{Float32,1}, h, bias,
                   iso::Bool, iters::Integer, act::Int32, creg, flags::Int32)
    d = make_desc(y, h, iters, iso, act, bias isa CuArray, flags, creg)
    first(forward!(d, y, h, λ, ρ, bias; ckpt=false))
end

function ChainRulesCore.rrule(::typeof(admm_call), y, λ, ρ, h, bias, iso, iters, act, creg, flags)
    d = make_desc(y, h, iters, iso, act, bias isa CuArray, flags, creg)
    x, ckb = forward!(d, y, h, λ, ρ, bias; ckpt=true)
    function admm_pullback(x̄)
        ȳ, h̄, λ̄, ρ̄, b̄ = backward!(d, CuArray{Float32,4}(unthunk(x̄)), x, y, h, λ, ρ, ckb)
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

# ---- the layers (deconv_admm.jl:6-212) --------------------------------------------------------------
for (T, trainables) in ((:ADMMDeconv, (:weight, :bias, :λ, :ρ)), (:ADMMDeconvF1, (:weight, :bias, :ρ)),
                        (:ADMMDeconvF2, (:weight, :bias, :λ)), (:ADMMDeconvF3, (:weight, :bias)))
    @eval begin
        mutable struct $T{F,A,N,V,M,B,C,D}
            σ::F
            weight::A
            bias::V
            λ::N
            ρ::M
            iters::B
            iso::C
            creg::D
        end
        # 8-positional form (w, σ, b, λ, ρ, iters, iso, creg), deconv_admm.jl:18-28,70-80,122-132,176-186
        $T(w::AbstractArray{T_,N_}, σ, b, lambda, rho, iters, iso, creg) where {T_,N_} =
            $T(σ, w, b, lambda, rho, iters, iso, creg)
        Flux.@layer $T trainable=$trainables
    end
end

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

end # module
