# DSMGPNative.jl -- ccall shim for libdsmgp.so (include/dsmgp.h).
#
# Written WITHOUT a Julia toolchain (none exists in the build image): it mirrors, call for call, the ctypes
# binding in deepstructuredmixtures_b200/_native.py + _handle.py, which IS exercised by the test-suite.
# Drop this file into DeepStructuredMixtures/src, `include("DSMGPNative.jl")` after "treeStructure.jl"
# (DeepStructuredMixtures.jl:134) and set ENV["DSMGP_LIB"] to the path of libdsmgp.so.
#
# The methods below REPLACE the bodies of the reference's hot-path methods; signatures are unchanged:
#   fit!(spn, D, gpmap; τ)            fit.jl:71        -> dsmgp_fit
#   mll!(spn, ℓ)                      optimize.jl:27   -> dsmgp_lml
#   updategradients!(spn) + ∇mll!     fit.jl:306, optimize.jl:42-150 -> dsmgp_grad
#   setparams!(spn, hyp)              optimize.jl:188  -> dsmgp_set_params
#   update!(spn)                      common.jl:326    -> dsmgp_update_weights
#   predict(model, x)                 common.jl:304-307-> dsmgp_predict
#   chol_continue!(A, ki)             AdvancedCholeskey.jl:152 -> dsmgp_chol_continue
module DSMGPNative

using ..DeepStructuredMixtures: GPNode, GPSplitNode, GPSumNode, DSMGP, PoE, gPoE, rBCM, IsoSE, ArdSE, IsoLinear,
    ArdLinear, KernelFunction, getLeaves, children
import SumProductNetworks: getOrderedNodes

const LIB = get(ENV, "DSMGP_LIB", "libdsmgp.so")

struct KernelDesc; type::Int32; nparams::Int32; end
struct Tree
    n_nodes::Int64; node_type::Ptr{Int32}; child_ptr::Ptr{Int64}; child_idx::Ptr{Int64}; leaf_of_node::Ptr{Int64}
    split_dim::Ptr{Int32}; split_ptr::Ptr{Int64}; split_val::Ptr{Float64}; root::Int64
end
mutable struct Opts
    as_written_grads::Int32; keep_factors::Int32; rank::Int32; world::Int32; device::Int32; strict_pd::Int32
    arena_bytes::Int64; reserved::NTuple{8,Int32}
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    nodeindex::Dict{Symbol,Int}      # node id -> 0-based node number (children before parents)
    nnodes::Int
    nparams::Int
end

check(rc, h=C_NULL) = rc == 0 ? nothing :
    error("libdsmgp error $rc: " * unsafe_string(ccall((:dsmgp_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

kerneltype(::IsoSE) = Int32(0); kerneltype(::ArdSE) = Int32(1)
kerneltype(::IsoLinear) = Int32(2); kerneltype(::ArdLinear) = Int32(3)
nkparams(k::KernelFunction) = Int32(length(k.logℓ) + 2)

"Flatten the region graph (post order: children before parents) into the arrays of `dsmgp_tree` + the leaf tables."
function flatten(spn)
    nodes = Any[]; index = Dict{Symbol,Int}()
    function rec(n)
        n isa GPNode || foreach(rec, children(n))
        index[n.id] = length(nodes); push!(nodes, n)
    end
    rec(spn)
    leaves = getLeaves(spn)
    leafno = Dict(l.id => i - 1 for (i, l) in enumerate(leaves))
    nn = length(nodes)
    node_type = zeros(Int32, nn); child_ptr = zeros(Int64, nn + 1); child_idx = Int64[]
    leaf_of_node = fill(Int64(-1), nn); split_dim = fill(Int32(-1), nn); split_ptr = zeros(Int64, nn + 1); split_val = Float64[]
    for (i, n) in enumerate(nodes)
        if n isa GPNode
            leaf_of_node[i] = leafno[n.id]
        else
            if n isa GPSplitNode
                node_type[i] = 1; split_dim[i] = n.split[1][1] - 1; append!(split_val, last.(n.split))
            else
                node_type[i] = eltype(children(n)) <: GPNode ? 3 : 2
            end
            append!(child_idx, [index[c.id] for c in children(n)])
        end
        child_ptr[i + 1] = length(child_idx); split_ptr[i + 1] = length(split_val)
    end
    isempty(child_idx) && push!(child_idx, 0); isempty(split_val) && push!(split_val, 0.0)
    leaf_ptr = vcat(0, cumsum([l.nobs for l in leaves]))
    leaf_obs = reduce(vcat, [Int64.(l.obs) for l in leaves])
    leaf_kid = Int32[l.kernelid - 1 for l in leaves]
    return (; nodes, index, leaves, nn, node_type, child_ptr, child_idx, leaf_of_node, split_dim, split_ptr, split_val,
            leaf_ptr, leaf_obs, leaf_kid, root = index[spn.id])
end

"getOverlap(spn, D, gpmap) fit.jl:12-39 on the device: the L×L overlap matrix (treeStructure.jl:428-431)."
function overlap(spn, N::Integer)
    f = flatten(spn); L = length(f.leaves); D = zeros(L, L)
    (; node_type, child_ptr, child_idx, leaf_of_node, split_dim, split_ptr, split_val) = f
    GC.@preserve node_type child_ptr child_idx leaf_of_node split_dim split_ptr split_val begin
        tree = Tree(f.nn, pointer(node_type), pointer(child_ptr), pointer(child_idx), pointer(leaf_of_node),
                    pointer(split_dim), pointer(split_ptr), pointer(split_val), f.root)
        check(ccall((:dsmgp_overlap, LIB), Int32,
                    (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ref{Tree}, Ptr{Float64}),
                    N, L, f.leaf_ptr, f.leaf_obs, f.leaf_kid, tree, D))
    end
    return D
end

"Create the device handle.  x is the GLOBAL N×D input matrix."
function create(spn, x::Matrix{Float64}; as_written=true, keep_factors=true)
    f = flatten(spn)
    (; index, leaves, nn, node_type, child_ptr, child_idx, leaf_of_node, split_dim, split_ptr, split_val,
       leaf_ptr, leaf_obs, leaf_kid) = f
    y_centered = reduce(vcat, [l.dist.y for l in leaves])              # already mean-subtracted (gaussianprocess.jl:72-74)
    leaf_mean = [l.dist.mean.m for l in leaves]
    nk = maximum(l.kernelid for l in leaves)
    kerns = [first(l for l in leaves if l.kernelid == k).dist.kernel for k in 1:nk]
    kd = [KernelDesc(kerneltype(k), nkparams(k)) for k in kerns]
    opts = Opts(0, 0, 0, 0, 0, 0, 0, ntuple(_ -> Int32(0), 8))
    ccall((:dsmgp_default_opts, LIB), Cvoid, (Ref{Opts},), opts)
    opts.as_written_grads = as_written ? 1 : 0; opts.keep_factors = keep_factors ? 1 : 0
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve node_type child_ptr child_idx leaf_of_node split_dim split_ptr split_val begin
        tree = Tree(nn, pointer(node_type), pointer(child_ptr), pointer(child_idx), pointer(leaf_of_node),
                    pointer(split_dim), pointer(split_ptr), pointer(split_val), index[spn.id])
        rc = ccall((:dsmgp_create, LIB), Int32,
                   (Ptr{Float64}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32},
                    Ptr{KernelDesc}, Int32, Ref{Tree}, Ref{Opts}, Ref{Ptr{Cvoid}}),
                   x, size(x, 1), size(x, 2), length(leaves), leaf_ptr, leaf_obs, y_centered, leaf_mean, leaf_kid,
                   kd, nk, tree, opts, out)
        check(rc)
    end
    h = Handle(out[], index, nn, Int(ccall((:dsmgp_nparams, LIB), Int64, (Ptr{Cvoid},), out[])))
    finalizer(h -> ccall((:dsmgp_destroy, LIB), Cvoid, (Ptr{Cvoid},), h.ptr), h)
    return h
end

setparams!(h::Handle, hyp::Vector{Float64}) =
    check(ccall((:dsmgp_set_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), h.ptr, hyp, length(hyp)), h.ptr)

"fit!(spn, D, gpmap; τ): the shared Cholesky with the overlap matrix D; returns elapsed seconds like fit.jl:88,121."
function fit!(h::Handle, D::Matrix{Float64}; τ::Float64 = 0.05)
    sec = Ref{Float64}(0.0)
    check(ccall((:dsmgp_fit, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Int32}, Ref{Float64}), h.ptr, τ, D, C_NULL, sec), h.ptr)
    return sec[]
end

"fit_naive!(spn) fit.jl:294-304: every expert factored on its own."
function fit_naive!(h::Handle)
    sec = Ref{Float64}(0.0)
    check(ccall((:dsmgp_fit, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Int32}, Ref{Float64}), h.ptr, 0.05, C_NULL, C_NULL, sec), h.ptr)
    return sec[]
end
fit!(h::Handle) = fit_naive!(h)

"mll!(spn, ℓ): fills the AxisArray keyed by node id (optimize.jl:27-39)."
function mll!(h::Handle, ℓ)
    tab = zeros(h.nnodes)
    check(ccall((:dsmgp_lml, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, tab), h.ptr)
    for (id, i) in h.nodeindex; ℓ[id] = tab[i + 1]; end
    return ℓ
end

"updategradients!(spn) followed by ∇mll!(spn, 0.0, 0.0, ℓ, ℓ[root], grad[, D[g,:], gpmap]); grad is overwritten."
function ∇mll!(h::Handle, grad::Vector{Float64}, Drow::Union{Nothing,Vector{Float64}}=nothing)
    check(ccall((:dsmgp_grad, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                h.ptr, Drow === nothing ? C_NULL : Drow, grad), h.ptr)
    return grad
end

"One train! iteration body (optimisers.jl:43-77 minus the Flux step)."
function evaluate!(h::Handle, hyp::Vector{Float64}, grad::Vector{Float64})
    lml = Ref{Float64}(0.0)
    check(ccall((:dsmgp_eval, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}),
                h.ptr, hyp, length(hyp), C_NULL, lml, grad, C_NULL), h.ptr)
    return lml[]
end

"update!(spn): writes node.logweights of every sum node, returns z (common.jl:326-332)."
function update!(h::Handle, spn)
    nodes = filter(n -> n isa GPSumNode, getOrderedNodes(spn))
    total = sum(length(children(n)) for n in getOrderedNodes(spn) if !(n isa GPNode))
    lw = zeros(max(total, 1)); z = Ref{Float64}(0.0)
    check(ccall((:dsmgp_update_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), h.ptr, lw, z), h.ptr)
    # lw is CSR by child_ptr in node-number order; recompute the offsets exactly like `create`
    off = 0; order = sort(collect(h.nodeindex), by = last)
    byid = Dict(n.id => n for n in getOrderedNodes(spn))
    for (id, _) in order
        n = byid[id]; n isa GPNode && continue
        k = length(children(n))
        n isa GPSumNode && (n.logweights[:] = lw[off + 1:off + k])
        off += k
    end
    return z[]
end

function _write_logweights!(h::Handle, spn, lw::Vector{Float64})
    off = 0; order = sort(collect(h.nodeindex), by = last)
    byid = Dict(n.id => n for n in getOrderedNodes(spn))
    for (id, _) in order
        n = byid[id]; n isa GPNode && continue
        k = length(children(n))
        n isa GPSumNode && (n.logweights[:] = lw[off + 1:off + k])
        off += k
    end
end

"infer!(spn) common.jl:336-355."
function infer!(h::Handle, spn)
    total = sum(length(children(n)) for n in getOrderedNodes(spn) if !(n isa GPNode))
    lw = zeros(max(total, 1)); z = Ref{Float64}(0.0)
    check(ccall((:dsmgp_infer, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), h.ptr, lw, z), h.ptr)
    _write_logweights!(h, spn, lw)
    return z[]
end

"reset_weights!(spn) common.jl:357-363."
function reset_weights!(h::Handle, spn)
    total = sum(length(children(n)) for n in getOrderedNodes(spn) if !(n isa GPNode))
    lw = zeros(max(total, 1))
    check(ccall((:dsmgp_reset_weights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, lw), h.ptr)
    _write_logweights!(h, spn, lw)
end

"""
Multi-GPU, one Julia process per GPU (handle created with rank/world): rank 0 calls `comm_unique_id()`, ships the 128 bytes to the
other ranks (MPI.jl bcast, a socket, a file), every rank calls `comm_init!(h, id)`.  After that fit! / evaluate! / update! / predict
are the same calls as on one GPU: the library all-reduces the per-leaf rows over NCCL itself.
"""
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:dsmgp_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    return id
end
comm_init!(h::Handle, id::Vector{UInt8}) = check(ccall((:dsmgp_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), h.ptr, id), h.ptr)

"""
train!(spn, D, gpmap, optim; iterations, λ, earlystop) optimisers.jl:40-83 as one call.  `optimiser`: 0 Descent(η), 1 ADAM(η, (β1, β2)),
2 RMSProp(η, ρ = β1); the reference's `hyp += grad` rebinding gives Flux a fresh state every iteration (state_by_identity).
Returns (final hyp, ℓ trace).
"""
function train!(h::Handle, hyp::Vector{Float64}; optimiser=1, η=0.001, β1=0.9, β2=0.999, state_by_identity=true,
                iterations=10_000, λ=0.05, earlystop=10)
    ℓ = zeros(iterations); nd = Ref{Int64}(0)
    check(ccall((:dsmgp_train, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Int32, Int64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
                h.ptr, optimiser, η, β1, β2, state_by_identity ? 1 : 0, iterations, λ, earlystop, hyp, ℓ, nd), h.ptr)
    return hyp, ℓ[1:nd[]]
end

"""
The inner loop of finetune! (finetuning.jl:36-58) for all anchor experts of one iteration in ONE call.
`anchors`: 0-based leaf numbers (gpmap order), `thetas`: H x G (one column per anchor), `D`: the L x L overlap matrix.
Returns (leaf_lml[G], grads H x G, root_lml[G]); the caller applies Flux.Optimise.apply! per anchor as before.
"""
function finetune_eval(h::Handle, anchors::Vector{Int64}, thetas::Matrix{Float64}, D::Matrix{Float64})
    G = length(anchors); H = size(thetas, 1)
    ll = zeros(G); gr = zeros(H, G); rl = zeros(G)       # column-major H x G == row-major G x H of the C side
    check(ccall((:dsmgp_finetune_eval, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                h.ptr, G, anchors, thetas, D, ll, gr, rl), h.ptr)
    return ll, gr, rl
end

predictmode(::DSMGP) = Int32(0); predictmode(::PoE) = Int32(1); predictmode(::gPoE) = Int32(2); predictmode(::rBCM) = Int32(3)
function predict(h::Handle, model, x::Matrix{Float64})
    T = size(x, 1); μ = zeros(T); σ² = zeros(T)
    check(ccall((:dsmgp_predict, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}),
                h.ptr, x, T, predictmode(model), μ, σ²), h.ptr)
    return μ, σ²
end

function chol_continue!(A::Matrix{Float64}, ki::Int)
    info = Ref{Int32}(0)
    check(ccall((:dsmgp_chol_continue, LIB), Int32, (Ptr{Float64}, Int64, Int64, Ref{Int32}), A, size(A, 1), ki, info))
    return LinearAlgebra.LowerTriangular(A), Int(info[])
end

# dsmgp_int8_info: what the last evaluation ran on the INT8 tensor cores (csrc/api_ozaki.cu; `ENV["DSMGP_OZAKI"] = "0"` before
# `create` keeps every flop on the FP64 pipelines).  Returns (batches, slices, int8_ops, fp64_equiv_flops, gemm_ms, slice_ms, fp64_tile_ms,
# pool_bytes, fp64_tile_flops).
function int8_info(h::Handle)
    o = zeros(Float64, 9)
    check(ccall((:dsmgp_int8_info, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), h.ptr, o, 9), h.ptr)
    return o
end

# dsmgp_host_split_plan (host-only): the diagonal ranges of an expert of n observations on the INT8 split path and the share of its
# factorisation + inverse flops that runs as INT8 block products.
function split_plan(n::Integer; depth::Integer=0, min_nb::Integer=0)
    nb = cld(cld(n, 64) * 64, 128)
    ro = zeros(Int32, nb); nr = Ref{Int32}(0); share = Ref{Float64}(0.0)
    rc = ccall((:dsmgp_host_split_plan, LIB), Int32, (Int64, Int32, Int32, Ptr{Int32}, Ref{Int32}, Ref{Float64}), n, depth, min_nb, ro, nr, share)
    rc == 0 || error("dsmgp_host_split_plan failed: $rc")
    return ro, Int(nr[]), share[]
end

end # module
