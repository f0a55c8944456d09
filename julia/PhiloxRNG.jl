# PhiloxRNG.jl -- the variate spec of libkissabc_cuda (DESIGN.md section 2) as a Julia AbstractRNG, so that the UNMODIFIED
# KissABC.jl can be run on exactly the variates the CPU oracle / the device consume:
#
#   * Philox4x32-10 word stream: key = 64-bit seed, counter = (block j, id, epoch, tag); words of block j, then block j+1, ...
#   * rand(rng)            u = (w + 0.5) 2^-32                         one word
#   * rand(rng, a:b)       a + floor(w n / 2^32), n = b - a + 1        one word     (also eachindex(v), tuples, ...)
#   * randn(rng)           Box-Muller, r = sqrt(-2 log u1), keeps r cos(2 pi u2)     two words
#   * randexp(rng)         -log u                                      one word
#   * spec_normals(rng,n)  what the registered simulators draw: words in groups of four -> 4 normals (both Box-Muller
#                          branches of two pairs); the unused normals of the last group are dropped
#   * log / exp / sincos(2 pi u) are the spec's fixed sequences of IEEE operations (spec_log, spec_exp, spec_sincos2pi):
#     the same bits as oracle/kabc_oracle.c and csrc/kabc_device.cuh on any IEEE machine.
#
# NOT executed in this repository's CI (no Julia in the build image): written against Julia >= 1.5 / Random stdlib.  The
# companion script make_ref_fixtures.jl runs KissABC.smc / KissABC.transition! on a PhiloxRNG in SERIAL mode (one stream,
# tag 6, consumed in the reference's own order) and writes tests/golden/ref_*.json; tests/test_ref_fixtures.py compares the
# oracle's serial mode with them.
module PhiloxRNGs

using Random
import Random: rand, randn, randexp, Sampler, SamplerTrivial, SamplerType, CloseOpen01, Repetition

export PhiloxRNG, spec_normals, spec_log, spec_exp, spec_sincos2pi, set_stream!, words_consumed

const ST_PRIOR, ST_PROPOSE, ST_COST, ST_ACCEPT, ST_COST_INIT, ST_SERIAL = UInt32(1), UInt32(2), UInt32(3), UInt32(4), UInt32(5), UInt32(6)

mutable struct PhiloxRNG <: AbstractRNG
    key::NTuple{2,UInt32}
    ctr::NTuple{4,UInt32}     # (block, id, epoch, tag); ctr[1] is the NEXT block to generate
    buf::NTuple{4,UInt32}
    pos::Int                  # words of buf already handed out (4 = empty)
    consumed::Int
end

PhiloxRNG(seed::UInt64; tag::UInt32=ST_SERIAL, id::Integer=0, epoch::Integer=0) =
    PhiloxRNG((UInt32(seed & 0xffffffff), UInt32(seed >> 32)), (UInt32(0), UInt32(id), UInt32(epoch), tag),
              (UInt32(0), UInt32(0), UInt32(0), UInt32(0)), 4, 0)

"reposition on the start of stream (tag, id, epoch)"
function set_stream!(r::PhiloxRNG, tag::UInt32, id::Integer, epoch::Integer)
    r.ctr = (UInt32(0), UInt32(id), UInt32(epoch), tag)
    r.pos = 4
    r
end
words_consumed(r::PhiloxRNG) = r.consumed

@inline function mulhilo(a::UInt32, b::UInt32)
    p = UInt64(a) * UInt64(b)
    UInt32(p >> 32), UInt32(p & 0xffffffff)
end

function philox4x32_10(ctr::NTuple{4,UInt32}, key::NTuple{2,UInt32})
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in 1:10
        hi0, lo0 = mulhilo(0xD2511F53, c0)
        hi1, lo1 = mulhilo(0xCD9E8D57, c2)
        c0, c1, c2, c3 = hi1 ⊻ c1 ⊻ k0, lo1, hi0 ⊻ c3 ⊻ k1, lo0
        k0 += 0x9E3779B9          # UInt32 arithmetic wraps
        k1 += 0xBB67AE85
    end
    (c0, c1, c2, c3)
end

function next_u32!(r::PhiloxRNG)
    if r.pos == 4
        r.buf = philox4x32_10(r.ctr, r.key)
        r.ctr = (r.ctr[1] + UInt32(1), r.ctr[2], r.ctr[3], r.ctr[4])
        r.pos = 0
    end
    r.pos += 1
    r.consumed += 1
    r.buf[r.pos]
end

# ---- the spec's elementary functions: fixed IEEE sequences (no fused contraction: Julia never contracts a*b+c by itself)
u01(w::UInt32) = (Float64(w) + 0.5) * 2.3283064365386962890625e-10

const LN2_HI = 6.93147180369123816490e-01
const LN2_LO = 1.90821492927058770002e-10
const LOG_C = (1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0)
function spec_log(x::Float64)
    isnan(x) && return x
    x < 0.0 && return NaN
    x == 0.0 && return -Inf
    x == Inf && return x
    e = 0
    b = reinterpret(UInt64, x)
    if (b >> 52) == 0
        x = x * 18014398509481984.0
        b = reinterpret(UInt64, x)
        e = -54
    end
    e += Int(b >> 52) - 1023
    m = reinterpret(Float64, (b & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000)
    if m > 1.4142135623730951
        m = m * 0.5
        e += 1
    end
    s = (m - 1.0) / (m + 1.0)
    s2 = s * s
    p = LOG_C[1]
    for q in 2:11
        p = fma(p, s2, LOG_C[q])
    end
    t = (s * s2) * p
    r = 2.0 * s + 2.0 * t
    ef = Float64(e)
    ef * LN2_HI + (r + ef * LN2_LO)
end

const EXP_C = (1.0 / 87178291200.0, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
               1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5)
function spec_exp(x::Float64)
    isnan(x) && return x
    x > 709.78 && return Inf
    x < -745.2 && return 0.0
    k = floor(x * 1.4426950408889634 + 0.5)
    r = fma(-k, LN2_HI, x)
    r = fma(-k, LN2_LO, r)
    p = EXP_C[1]
    for q in 2:13
        p = fma(p, r, EXP_C[q])
    end
    p = fma(p, r, 1.0)
    p = fma(p, r, 1.0)
    ki = Int(k)
    k1 = div(ki, 2)            # truncating division, as in C
    k2 = ki - k1
    f1 = reinterpret(Float64, UInt64(k1 + 1023) << 52)
    f2 = reinterpret(Float64, UInt64(k2 + 1023) << 52)
    (p * f1) * f2
end

const SIN_C = (-1.0 / 121645100408832000.0, 1.0 / 355687428096000.0, -1.0 / 1307674368000.0, 1.0 / 6227020800.0,
               -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0)
const COS_C = (1.0 / 6402373705728000.0, -1.0 / 20922789888000.0, 1.0 / 87178291200.0, -1.0 / 479001600.0, 1.0 / 3628800.0,
               -1.0 / 40320.0, 1.0 / 720.0, -1.0 / 24.0, 0.5)
"(sin 2 pi u, cos 2 pi u)"
function spec_sincos2pi(u::Float64)
    q = floor(4.0 * u + 0.5)
    t = u - 0.25 * q
    phi = t * 6.283185307179586
    p2 = phi * phi
    ps, pc = SIN_C[1], COS_C[1]
    for j in 2:9
        ps = fma(ps, p2, SIN_C[j])
        pc = fma(pc, p2, COS_C[j])
    end
    s = fma(phi * p2, ps, phi)
    c = fma(-p2, pc, 1.0)
    qi = Int(q) & 3
    sn = qi == 0 ? s : (qi == 1 ? c : (qi == 2 ? -s : -c))
    cs = qi == 0 ? c : (qi == 1 ? -s : (qi == 2 ? -c : s))
    sn, cs
end

function normal_pair(w0::UInt32, w1::UInt32)
    r = sqrt(-2.0 * spec_log(u01(w0)))
    s, c = spec_sincos2pi(u01(w1))
    r * c, r * s
end

"n normals the way the registered simulators draw them: 4 words -> 4 normals, leftovers of the last group dropped"
function spec_normals(r::PhiloxRNG, n::Integer)
    z = Vector{Float64}(undef, n)
    j = 0
    while j < n
        w0 = next_u32!(r); w1 = next_u32!(r); w2 = next_u32!(r); w3 = next_u32!(r)
        a, b = normal_pair(w0, w1)
        c, d = normal_pair(w2, w3)
        for v in (a, b, c, d)
            j < n && (z[j += 1] = v)
        end
    end
    z
end

# ---- Random API
rand(r::PhiloxRNG, ::SamplerType{UInt32}) = next_u32!(r)
rand(r::PhiloxRNG, ::SamplerType{UInt64}) = (UInt64(next_u32!(r)) << 32) | UInt64(next_u32!(r))
rand(r::PhiloxRNG, ::SamplerTrivial{CloseOpen01{Float64}}) = u01(next_u32!(r))

# every integer range (1:n, eachindex(v), Base.OneTo(n) behind tuples / arrays) : first + floor(w n / 2^32)
struct SpecRange{T<:Integer} <: Sampler{T}
    first::T
    n::UInt32
end
Sampler(::Type{PhiloxRNG}, rg::AbstractUnitRange{T}, ::Repetition) where {T<:Base.BitInteger} = SpecRange{T}(first(rg), UInt32(length(rg)))
rand(r::PhiloxRNG, sp::SpecRange{T}) where {T} = sp.first + T((UInt64(next_u32!(r)) * UInt64(sp.n)) >> 32)
# the two call shapes the reference uses, routed directly (independent of the Sampler internals of the Julia version):
# rand(rng, 1:n) / rand(rng, eachindex(v))  (src/smc.jl:163-164, src/transition.jl:6-54)  and  rand(rng, (1,1,1,1,2,2,3))  (:62)
rand(r::PhiloxRNG, rg::AbstractUnitRange{T}) where {T<:Integer} = first(rg) + T((UInt64(next_u32!(r)) * UInt64(length(rg))) >> 32)
rand(r::PhiloxRNG, t::Tuple) = t[rand(r, 1:length(t))]
# a tuple of Ints is also a `Dims`: Random has `rand(::AbstractRNG, ::Dims)` ("needed to disambiguate"), which would make the
# call above ambiguous for exactly the tuple the reference draws from -- settle it with the most specific method
rand(r::PhiloxRNG, t::Dims) = t[rand(r, 1:length(t))]

function randn(r::PhiloxRNG)
    w0 = next_u32!(r); w1 = next_u32!(r)
    normal_pair(w0, w1)[1]
end
randn(r::PhiloxRNG, ::Type{Float64}) = randn(r)
randexp(r::PhiloxRNG) = -spec_log(u01(next_u32!(r)))
randexp(r::PhiloxRNG, ::Type{Float64}) = randexp(r)

end # module
