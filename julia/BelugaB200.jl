# BelugaB200.jl — Julia side of the drop-in: the methods of Beluga.jl's likelihood hot path,
# forwarded through `ccall` to libbeluga_b200.so (include/beluga_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same ABI is exercised
# call-for-call by the Python mirror (beluga.jl_b200/profile.py, whose `sync` this file's `sync!`
# follows line by line) in tests/test_gpu_parity.py.
#
# Usage: `include("BelugaB200.jl")` inside module Beluga, after src/profile.jl.  Nothing else in
# Beluga.jl is edited — every signature of the public API stays:
#   DLWGD(nw, df, λ, μ, η, nt)            src/model.jl:22-28      unchanged: it reaches the profile through
#   Profile(X, n), PArray()               src/profile.jl:28,36-37 -> DeviceProfile (PArray(): the mock, no GPU)
#   logpdf!(d, p), logpdf!(n, p)          src/profile.jl:56-60
#   logpdfroot(n, p)                      src/profile.jl:62-63
#   gradient(d, p), gradient_cr(d, p)     src/profile.jl:71-75
#   set!(p), rev!(p)                      src/profile.jl:79-80
#   extend!(p, i), shrink!(p, i)          src/profile.jl:108-109
#   set!(n::ModelNode)                    src/node.jl:118-121     -> records the request (ε, W are built
#                                                                    on the device at the next likelihood call)
# rjmcmc.jl, mcmc.jl, priors.jl are untouched: their closures call `logpdf!(n, data)` exactly as before;
# the model is found from the node itself (walk to the root, then `prewalk`).

const LIBBLG = get(ENV, "BELUGA_B200_LIB", "libbeluga_b200.so")

mutable struct DeviceProfile{T<:Real}          # stands in for PArray{T}
    h::Ptr{Cvoid}                               # C_NULL: the mock profile PArray() (profile.jl:28,31)
    npatterns::Int
    # what the device holds (sync! uploads only what differs)
    parent::Vector{Int32}
    kind::Vector{UInt8}
    cpos::Vector{Int32}
    θ::Matrix{Float64}                          # N × 4: λ, μ, q, t by id
    η::Float64
    mk::Int
    built::Bool                                 # a full ε/W build has been done for this topology
    function DeviceProfile{T}(h, n) where T
        p = new{T}(h, n, Int32[], UInt8[], Int32[], zeros(0, 4), NaN, -1, false)
        h == C_NULL || finalizer(x -> ccall((:blg_destroy, LIBBLG), Cint, (Ptr{Cvoid},), x.h), p)
        p
    end
end
DeviceProfile() = DeviceProfile{Float64}(C_NULL, 0)          # PArray()
Base.length(p::DeviceProfile) = p.npatterns

blgcheck(p, rc) = rc == 0 || error(unsafe_string(ccall((:blg_last_error, LIBBLG), Cstring, (Ptr{Cvoid},), p.h)))

# Profile(X, n) — src/profile.jl:36-37.  X = permutedims(M) (nodes × patterns), n = multiplicities.
function DeviceProfile(X::Matrix{Int64}, n::Vector{Int64}; device::Int=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:blg_create, LIBBLG), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), device, C_NULL, h)
    rc == 0 || error(unsafe_string(ccall((:blg_last_error, LIBBLG), Cstring, (Ptr{Cvoid},), C_NULL)))
    p = DeviceProfile{Float64}(h[], length(n))
    X32 = Int32.(X)                                              # column-major nodes×patterns == node fastest
    blgcheck(p, ccall((:blg_set_profiles, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}, Cint, Cint),
                      p.h, X32, n, size(X, 1), size(X, 2)))
    p
end

# The two constructors the rest of Beluga.jl reaches profiles through (src/profile.jl:28,36-37).  With them replaced,
# `profile(t, l, df)` (src/model.jl:40-49) and hence `DLWGD(nw, df, λ, μ, η, nt)` (src/model.jl:22-28) return the device
# profile WITHOUT being edited, and `DLWGD(nw, λ, μ, η, nt)` (src/model.jl:30-36) returns the mock.
Profile(X::Matrix{Int64}, n::Vector{Int64}) = DeviceProfile(X, n; device=parse(Int, get(ENV, "BELUGA_B200_DEVICE", "0")))
PArray() = DeviceProfile()

# ---- set!(n): record instead of compute ------------------------------------------------------------------------
# update!(n, …) (node.jl:99-116) ends in setabove!/setbelow!, which call set!(n) node by node.  ε and W are not
# needed on the host any more (priors and moves never read n.x.W or n.x.ϵ), so set!(n) only records the id; the next
# likelihood call replays the recorded ids, in order, on the device (blg_rebuild).
const SET_REQUESTS = Int32[]
const FULL_REQUEST = Ref(false)
set!(n::ModelNode) = (push!(SET_REQUESTS, Int32(n.i)); n)        # replaces node.jl:118-121
set!(d::DLWGD) = (FULL_REQUEST[] = true; d)                      # replaces model.jl:56-60 (set!(d): every node)

kindcode(n) = n.x.kind == :root ? 0x00 : n.x.kind == :sp ? 0x01 : n.x.kind == :wgd ? 0x02 :
              n.x.kind == :wgt ? 0x03 : 0x04

rootof(n::ModelNode) = (while !isroot(n); n = n.p; end; n)
nodesof(r::ModelNode) = Dict(n.i => n for n in prewalk(r))

# mirror topology + θ and replay the recorded set!(n) calls on the device — profile.py::PArray.sync, 1:1
function sync!(p::DeviceProfile, nodes::Dict)
    N = maximum(keys(nodes))
    parent = zeros(Int32, N); kind = fill(0xff, N); cpos = zeros(Int32, N)
    θ = zeros(N, 4); η = NaN
    for (i, n) in nodes
        kind[i] = kindcode(n)
        if !isroot(n)
            parent[i] = n.p.i
            cpos[i] = findfirst(c -> c === n, collect(n.p.c)) - 1
        else
            η = n.x.θ[:η]
        end
        x = n.x.θ
        θ[i, 1] = get(x, :λ, 0.); θ[i, 2] = get(x, :μ, 0.); θ[i, 3] = get(x, :q, 0.); θ[i, 4] = get(x, :t, 0.)
    end
    mk = nodetype(nodes[1]) == Node ? 0 : 1
    full = FULL_REQUEST[]
    if parent != p.parent || kind != p.kind || cpos != p.cpos       # addwgd!/removewgd!/reindex!, or a new model
        blgcheck(p, ccall((:blg_set_topology, LIBBLG), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{UInt8}, Ptr{Int32}),
                          p.h, N, parent, kind, cpos))
        p.parent, p.kind, p.cpos = parent, kind, cpos
        p.θ = zeros(0, 4); full = true
    end
    if θ != p.θ || η != p.η || mk != p.mk
        blgcheck(p, ccall((:blg_set_params, LIBBLG), Cint,
                          (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64),
                          p.h, mk, θ[:, 1], θ[:, 2], θ[:, 3], θ[:, 4], η))
        p.θ, p.η, p.mk = θ, η, mk
    end
    if full || !p.built
        blgcheck(p, ccall((:blg_rebuild, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint), p.h, C_NULL, 0))
        p.built = true
    elseif !isempty(SET_REQUESTS)                                  # only the recorded path (setabove!/setbelow!)
        blgcheck(p, ccall((:blg_rebuild, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint), p.h, SET_REQUESTS, length(SET_REQUESTS)))
    end
    empty!(SET_REQUESTS); FULL_REQUEST[] = false
end
sync!(p::DeviceProfile, d::DLWGD) = sync!(p, d.nodes)

# src/profile.jl:56-63 — same signatures; with a communicator (comm_init!) the value is the sum over all ranks
function logpdf!(d::DLWGD, p::DeviceProfile{T}) where T
    p.h == C_NULL && return zero(T)                                # profile.jl:56 (mock / empty profile)
    sync!(p, d); out = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_logpdf_full, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, out)); out[]
end
function logpdf!(n::ModelNode, p::DeviceProfile{T}) where T
    p.h == C_NULL && return zero(T)
    sync!(p, nodesof(rootof(n))); out = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_logpdf_from, LIBBLG), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), p.h, n.i, out)); out[]
end
function logpdfroot(n::ModelNode, p::DeviceProfile{T}) where T
    p.h == C_NULL && return zero(T)
    sync!(p, nodesof(rootof(n))); out = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_logpdf_root, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, out)); out[]
end

# src/profile.jl:71-75.  The library orders the q block by ASCENDING wgd id; asvector (src/model.jl:541-546)
# lists q in getwgds order (descending).  reverse! restores the reference's layout.
function gradient(d::DLWGD, p::DeviceProfile)
    n = Beluga.ne(d) + 1; k = Beluga.nwgd(d)
    g = zeros(2n + k + 1); l = Ref{Float64}(0.)
    p.h == C_NULL && return g
    sync!(p, d)
    blgcheck(p, ccall((:blg_gradient, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}), p.h, g, length(g), l))
    reverse!(view(g, 2n+1:2n+k)); g
end
function gradient_cr(d::DLWGD, p::DeviceProfile)
    g = zeros(2); l = Ref{Float64}(0.)
    p.h == C_NULL && return g
    sync!(p, d)
    blgcheck(p, ccall((:blg_gradient_cr, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.h, g, l)); g
end

# src/profile.jl:79-80, 108-109 (no-ops on the mock profile, like the reference's loops over zero families)
set!(p::DeviceProfile) = (p.h == C_NULL || blgcheck(p, ccall((:blg_commit, LIBBLG), Cint, (Ptr{Cvoid},), p.h)); p)
rev!(p::DeviceProfile) = (p.h == C_NULL || blgcheck(p, ccall((:blg_revert, LIBBLG), Cint, (Ptr{Cvoid},), p.h)); p)
extend!(p::DeviceProfile, i::Int64) = (p.h == C_NULL || blgcheck(p, ccall((:blg_extend, LIBBLG), Cint, (Ptr{Cvoid}, Cint), p.h, i)); p)
shrink!(p::DeviceProfile, i::Int64) = (p.h == C_NULL || blgcheck(p, ccall((:blg_shrink, LIBBLG), Cint, (Ptr{Cvoid}, Cint), p.h, i)); p)

# ---- multi-GPU: one Julia process per GPU (the reference's Distributed workers), the library owns NCCL ---------------
# rank 0:  id = comm_unique_id();  every rank receives `id` (MPI.Bcast! / remotecall) and calls
# comm_init!(data, nranks, rank, id) after building its DeviceProfile from its contiguous block of patterns
# (profile.jl:27-37 distributes one PArray the same way).  From then on logpdf!/gradient return all-reduced values.
function comm_unique_id()
    id = zeros(UInt8, 128)
    ccall((:blg_comm_unique_id, LIBBLG), Cint, (Ptr{UInt8},), id) == 0 || error("NCCL not available")
    id
end
comm_init!(p::DeviceProfile, nranks::Int, rank::Int, id::Vector{UInt8}) =
    (blgcheck(p, ccall((:blg_comm_init, LIBBLG), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), p.h, nranks, rank, id)); p)
