# BelugaB200.jl — Julia side of the drop-in: the methods of Beluga.jl's likelihood hot path,
# forwarded through `ccall` to libbeluga_b200.so (include/beluga_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same ABI is exercised
# call-for-call by the Python mirror (beluga.jl_b200/profile.py) in tests/test_gpu_parity.py.
#
# Usage inside Beluga.jl: include this file after src/profile.jl and the methods below replace
#   Profile/PArray construction           src/profile.jl:27-37
#   logpdf!(d, p), logpdf!(n, p)          src/profile.jl:56-60
#   logpdfroot(n, p)                      src/profile.jl:62-63
#   gradient(d, p), gradient_cr(d, p)     src/profile.jl:71-75
#   set!(p), rev!(p)                      src/profile.jl:79-80
#   extend!(p, i), shrink!(p, i)          src/profile.jl:108-109
# The model graph (DLWGD, TreeNode, update!, addwgd!, removewgd!, …) is untouched; ε and W are no
# longer needed on the host (src/node.jl:136-218), so `set!(n::ModelNode)` may become a no-op once
# every consumer goes through the device (priors and moves never read n.x.W or n.x.ϵ).

const LIBBLG = get(ENV, "BELUGA_B200_LIB", "libbeluga_b200.so")

mutable struct DeviceProfile{T<:Real}          # stands in for PArray{T}
    h::Ptr{Cvoid}
    npatterns::Int
    model_id::UInt                              # objectid of the model last mirrored
    log::Vector{Vector{Int32}}                  # set!(n) requests recorded since the last sync
    function DeviceProfile{T}(h, n) where T
        p = new{T}(h, n, 0, Vector{Int32}[])
        finalizer(x -> ccall((:blg_destroy, LIBBLG), Cint, (Ptr{Cvoid},), x.h), p)
        p
    end
end

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

kindcode(n) = n.x.kind == :root ? 0x00 : n.x.kind == :sp ? 0x01 : n.x.kind == :wgd ? 0x02 :
              n.x.kind == :wgt ? 0x03 : 0x04

# mirror topology + θ and replay the reference's set!(n) calls (src/node.jl:99-134) on the device
function sync!(p::DeviceProfile, d::DLWGD)
    N = maximum(keys(d.nodes))
    parent = zeros(Int32, N); kind = fill(0xff, N); cpos = zeros(Int32, N)
    λ = zeros(N); μ = zeros(N); q = zeros(N); t = zeros(N)
    for (i, n) in d.nodes
        kind[i] = kindcode(n)
        if !isroot(n)
            parent[i] = n.p.i
            cpos[i] = findfirst(c -> c === n, collect(n.p.c)) - 1
        end
        θ = n.x.θ
        λ[i] = get(θ, :λ, 0.); μ[i] = get(θ, :μ, 0.); q[i] = get(θ, :q, 0.); t[i] = get(θ, :t, 0.)
    end
    mk = nodetype(d[1]) == Node ? 0 : 1
    blgcheck(p, ccall((:blg_set_topology, LIBBLG), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{UInt8}, Ptr{Int32}),
                      p.h, N, parent, kind, cpos))
    blgcheck(p, ccall((:blg_set_params, LIBBLG), Cint,
                      (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64),
                      p.h, mk, λ, μ, q, t, d[1][:η]))
    # simplest faithful policy: full on-device rebuild (set!(d), src/model.jl:56-60).  A production glue keeps
    # the topology resident and passes only the path recorded by update! (see beluga.jl_b200/profile.py::sync).
    blgcheck(p, ccall((:blg_rebuild, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint), p.h, C_NULL, 0))
end

# src/profile.jl:56-63
function logpdf!(d::DLWGD, p::DeviceProfile{T}) where T
    p.npatterns == 0 && return zero(T)
    sync!(p, d); out = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_logpdf_full, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, out)); out[]
end
function logpdf!(n::ModelNode, p::DeviceProfile{T}, d::DLWGD) where T
    p.npatterns == 0 && return zero(T)
    sync!(p, d); out = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_logpdf_from, LIBBLG), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), p.h, n.i, out)); out[]
end
function logpdfroot(n::ModelNode, p::DeviceProfile{T}, d::DLWGD) where T
    p.npatterns == 0 && return zero(T)
    sync!(p, d); out = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_logpdf_root, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, out)); out[]
end

# src/profile.jl:71-75.  The library orders the q block by ASCENDING wgd id; asvector (src/model.jl:541-546)
# lists q in getwgds order (descending).  reverse! restores the reference's layout.
function gradient(d::DLWGD, p::DeviceProfile)
    sync!(p, d)
    n = Beluga.ne(d) + 1; k = Beluga.nwgd(d)
    g = zeros(2n + k + 1); l = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_gradient, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}), p.h, g, length(g), l))
    reverse!(view(g, 2n+1:2n+k)); g
end
function gradient_cr(d::DLWGD, p::DeviceProfile)
    sync!(p, d); g = zeros(2); l = Ref{Float64}(0.)
    blgcheck(p, ccall((:blg_gradient_cr, LIBBLG), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.h, g, l)); g
end

# src/profile.jl:79-80, 108-109
set!(p::DeviceProfile) = (blgcheck(p, ccall((:blg_commit, LIBBLG), Cint, (Ptr{Cvoid},), p.h)); p)
rev!(p::DeviceProfile) = (blgcheck(p, ccall((:blg_revert, LIBBLG), Cint, (Ptr{Cvoid},), p.h)); p)
extend!(p::DeviceProfile, i::Int64) = (blgcheck(p, ccall((:blg_extend, LIBBLG), Cint, (Ptr{Cvoid}, Cint), p.h, i)); p)
shrink!(p::DeviceProfile, i::Int64) = (blgcheck(p, ccall((:blg_shrink, LIBBLG), Cint, (Ptr{Cvoid}, Cint), p.h, i)); p)
