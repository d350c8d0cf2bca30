# PTMCB200.jl — the binding a maintainer of ParallelTemperingMonteCarlo.jl adds to run the sweep hot path on a B200
# through libptmc_b200.so (C ABI: include/ptmc_b200.h).  Thin by design: it converts the package's own types into the
# POD config + flat arrays of the ABI, calls `ccall`, and fills the package's own `MCState`s / `Output` back.
#
# NOT executed in the build image (no Julia there); the same call sequence is exercised from Python by
# paralleltemperingmontecarlo.jl_b200/engine.py + host.py (ptmc_run), which the tests and bench.py drive.
#
#   using ParallelTemperingMonteCarlo, PTMCB200
#   backend = B200Backend(lib = "/path/to/libptmc_b200.so", device = 0, n_ladders = 1)
#   states, results = ptmc_run!(backend, mc_params, temp, start_config, potential, ensemble)
module PTMCB200

using ParallelTemperingMonteCarlo
using ParallelTemperingMonteCarlo: MCParams, TempGrid, Config, MCState, Output, NVT, NPT,
    SphericalBC, CubicBC, RhombicBC, RectangularBC, ELJPotential, ELJPotentialEven, EmbeddedAtomPotential,
    initialisation, finalise_results, initialise_histograms!, check_e_bounds, get_volume

export B200Backend, ptmc_run!

const PTMC_ABI_VERSION = Int32(3)
const PTMC_MAX_COEFF = 16
const PTMC_NCOUNT = 10

Base.@kwdef struct B200Backend
    lib::String
    device::Int = 0
    n_ladders::Int = 1          # independent ladders (seeds) batched on the device; results of ladder 1 are returned
    seed::UInt64 = 1234
    rng_mode::Int32 = 2         # PTMC_RNG_BATCH8
    lanes_per_replica::Int32 = 0
    chunk::Int = 100_000        # cycles per ccall (the driver touches states only at save / saveconfigs cadence)
end

# mirror of `ptmc_config` (include/ptmc_b200.h); field order and types must match the header
struct PtmcConfig
    abi_version::Int32
    device::Int32
    n_ladders::Int64
    n_traj::Int32
    n_atoms::Int32
    ensemble::Int32
    n_atom_moves::Int32
    n_volume_moves::Int32
    bc_kind::Int32
    pressure::Float64
    pot_kind::Int32
    n_coeff::Int32
    coeff::NTuple{PTMC_MAX_COEFF,Float64}
    eam_n::Float64
    eam_m::Float64
    eam_ean::Float64
    eam_ecam::Float64
    runner::Ptr{Cvoid}
    betas::Ptr{Float64}
    n_adjust::Int32
    mc_sample::Int32
    n_bin::Int32
    rng_mode::Int32
    min_acc::Float64
    max_acc::Float64
    seed::UInt64
    lanes_per_replica::Int32
    separated_volume::Int32
    eljb_a::Ptr{Float64}       # ABI 2: ELJPotentialB coeff_a / coeff_b (coeff_c travels in coeff)
    eljb_b::Ptr{Float64}
    lut_table::Ptr{Float64}    # ABI 3: LookupTablePotential (table transposed to [l_angle][l_dist] row-major = Julia's table' column-major)
    lut_l_dist::Int32
    lut_l_angle::Int32
    lut_start_dist::Float64
    lut_d_dist::Float64
    lut_start_angle::Float64
    lut_d_angle::Float64
    lut_c6coeff::Float64
end

bc_kind(::SphericalBC) = Int32(0); bc_kind(::CubicBC) = Int32(1); bc_kind(::RhombicBC) = Int32(2); bc_kind(::RectangularBC) = Int32(3)
box_of(bc::SphericalBC) = (bc.radius2, 0.0)
box_of(bc::CubicBC) = (bc.box_length, bc.box_length)
box_of(bc::Union{RhombicBC,RectangularBC}) = (bc.box_length, bc.box_height)

pot_fields(p::ELJPotential{N}) where {N} = (Int32(0), Int32(N), Tuple(p.coeff), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::ELJPotentialEven{N}) where {N} = (Int32(1), Int32(N), Tuple(p.coeff), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::EmbeddedAtomPotential) = (Int32(2), Int32(0), (), (p.n, p.m, p.ean, p.eCam))
pot_fields(p::LookupTablePotential) = (Int32(5), Int32(0), (), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::ELJPotentialB{N}) where {N} = (Int32(4), Int32(N), Tuple(p.coeff_c), (0.0, 0.0, 0.0, 0.0))

pot_fields(p::RuNNerPotential) = (Int32(3), Int32(0), (), (0.0, 0.0, 0.0, 0.0))

# mirror of `ptmc_runner_desc` (include/ptmc_b200.h); field order and types must match the header
struct PtmcRunnerDesc
    n_rad::Int32
    n_ang::Int32
    rad_eta::Ptr{Float64}
    rad_type::Ptr{Int32}
    rad_gmax_gmin::Ptr{Float64}
    ang_eta::Ptr{Float64}
    ang_lambda::Ptr{Float64}
    ang_zeta::Ptr{Float64}
    ang_type::Ptr{Int32}
    ang_gmax_gmin::Ptr{Float64}
    r_cut::Float64
    boundary2::Float64
    n_layers::Int32
    num_nodes::Ptr{Int32}
    activation::Ptr{Int32}
    n_params::Int32
    params::Ptr{Float64}
end

# RuNNerPotential -> descriptor + the arrays it points into (keep them alive across ptmc_create).  The package's symmetry-function
# structs hold G_norm = 1 / (G_max - G_min) and G_offset = -G_min * G_norm (SymmFunc.jl:40-44,107-112), the descriptor wants the
# scaling.data pair itself: a script that still has its `G_value_vec` should pass that instead (`scaling = G_value_vec`), because
# the pair recovered here, G_min = -G_offset / G_norm and G_max = G_min + 1 / G_norm, can differ from the file's by an ulp.
function runner_desc(p::RuNNerPotential{Nrad,Nang}; scaling=nothing) where {Nrad,Nang}
    rs, as = p.radsymfunctions, p.angsymfunctions
    gpair(norm, off) = (gmin = -off / norm; (gmin + 1 / norm, gmin))
    rg = Float64[]; ag = Float64[]
    for k in 1:Nrad
        append!(rg, scaling === nothing ? gpair(rs.G_norm[k], rs.G_offset[k]) : (scaling[k][1], scaling[k][2]))
    end
    for k in 1:Nang
        norm = as.tpz[k] / 2.0^(1 - as.zeta[k])
        append!(ag, scaling === nothing ? gpair(norm, as.G_offset[k]) : (scaling[Nrad + k][1], scaling[Nrad + k][2]))
    end
    arrays = (collect(Float64, rs.eta), collect(Int32, rs.type_vec), rg, collect(Float64, as.eta), collect(Float64, as.lambda),
              collect(Float64, as.zeta), collect(Int32, as.type_vec), ag, collect(Int32, p.nnp.num_nodes),
              collect(Int32, p.nnp.activation_functions), collect(Float64, p.nnp.parameters))
    desc = PtmcRunnerDesc(Nrad, Nang, pointer(arrays[1]), pointer(arrays[2]), pointer(arrays[3]), pointer(arrays[4]), pointer(arrays[5]),
                          pointer(arrays[6]), pointer(arrays[7]), pointer(arrays[8]), p.r_cut, p.boundary, p.nnp.n_layers,
                          pointer(arrays[9]), pointer(arrays[10]), p.nnp.n_params, pointer(arrays[11]))
    return Ref(desc), arrays
end

pad16(t) = ntuple(i -> i <= length(t) ? Float64(t[i]) : 0.0, PTMC_MAX_COEFF)

function check(lib, h, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:ptmc_last_error, lib), Cstring, (Ptr{Cvoid},), h))
    error("libptmc_b200: $msg (code $rc)")     # same text as the reference for PTMC_ERR_SPHERE
end

"""
    ptmc_run!(backend::B200Backend, mc_params, temp, start_config, potential, ensemble; save=false, saveconfigs=false, ...)

Drop-in for `ptmc_run!` (src/MCRun.jl:363-432): same arguments after `backend`, same return value
`(mc_states, results)`.  The sweep (`mc_cycle!`, src/MCRun.jl:192-253) runs on the device in chunks; states come
back to the host only where the reference's driver touches them (histogram bounds after equilibration, checkpoints,
end of run).
"""
function ParallelTemperingMonteCarlo.ptmc_run!(b::B200Backend, mc_params::MCParams, temp::TempGrid, start_config,
                                               potential, ensemble; rdfsave=false, save=false, saveconfigs=false, configsname="configuration", kwargs...)
    lib = b.lib
    # the package's own initialisation builds MCStates (host copies), MoveStrategy and the empty Output
    mc_states, move_strategy, results, n_steps, _ = initialisation(mc_params, temp, start_config, potential, ensemble)
    T, N, L = mc_params.n_traj, mc_params.n_atoms, b.n_ladders
    bc = mc_states[1].config.boundary_condition
    pk, nc, coeff, eam = pot_fields(potential)
    npt = ensemble isa NPT
    betas = collect(Float64, temp.beta_grid)
    ea = potential isa ELJPotentialB ? collect(Float64, potential.coeff_a) : Float64[]
    eb = potential isa ELJPotentialB ? collect(Float64, potential.coeff_b) : Float64[]
    islut = potential isa LookupTablePotential
    lut = islut ? collect(Float64, permutedims(potential.table)) : Float64[]        # memory order: distance index fastest
    rdesc, rkeep = potential isa RuNNerPotential ? runner_desc(potential; scaling=get(kwargs, :scaling, nothing)) : (nothing, ())
    cfg = Ref(PtmcConfig(PTMC_ABI_VERSION, b.device, L, T, N, npt ? 1 : 0, ensemble.n_atom_moves,
        npt ? ensemble.n_volume_moves : 0, bc_kind(bc), npt ? ensemble.pressure : 0.0, pk, nc, pad16(coeff),
        eam..., rdesc === nothing ? Ptr{Cvoid}(C_NULL) : Ptr{Cvoid}(Base.unsafe_convert(Ptr{PtmcRunnerDesc}, rdesc)), pointer(betas), mc_params.n_adjust, mc_params.mc_sample, mc_params.n_bin, b.rng_mode,
        mc_params.min_acc, mc_params.max_acc, b.seed, b.lanes_per_replica, (npt && ensemble.separated_volume) ? 1 : 0,
        isempty(ea) ? Ptr{Float64}(C_NULL) : pointer(ea), isempty(eb) ? Ptr{Float64}(C_NULL) : pointer(eb),
        islut ? pointer(lut) : Ptr{Float64}(C_NULL), islut ? potential.l_dist : Int32(0), islut ? potential.l_angle : Int32(0),
        islut ? potential.start_dist : 0.0, islut ? potential.d_dist : 0.0, islut ? potential.start_angle : 0.0,
        islut ? potential.d_angle : 0.0, islut ? potential.c6coeff : 0.0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve betas ea eb lut rdesc rkeep begin
        rc = ccall((:ptmc_create, lib), Cint, (Ref{PtmcConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    end
    rc == 0 || error("libptmc_b200: " * unsafe_string(ccall((:ptmc_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    try
        # state in: positions [L][T][N][3] (row-major = Julia (3, N, T, L)), box [L][T][2]
        pos = Array{Float64}(undef, 3, N, T, L); box = Array{Float64}(undef, 2, T, L)
        for l in 1:L, t in 1:T
            for (j, p) in enumerate(mc_states[t].config.positions); pos[:, j, t, l] .= p; end
            box[:, t, l] .= box_of(mc_states[t].config.boundary_condition)
        end
        check(lib, h[], ccall((:ptmc_upload_states, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h[], pos, box, C_NULL))
        # equilibration_cycle! (src/MCRun.jl:282-313): MC for eq_cycles ÷ 2 cycles, then the energy bounds
        n_eq = mc_params.eq_cycles ÷ 2
        n_eq > 0 && check(lib, h[], ccall((:ptmc_run_cycles, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Cint), h[], 1, n_eq, 0))
        en = Array{Float64}(undef, T, L)
        check(lib, h[], ccall((:ptmc_download_states, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                              h[], C_NULL, C_NULL, en, C_NULL, C_NULL, C_NULL))
        ebounds = [100.0, -100.0]
        for _ in (n_eq + 1):mc_params.eq_cycles, t in 1:T
            ebounds = check_e_bounds(en[t, 1], ebounds)
        end
        initialise_histograms!(mc_params, results, ebounds, bc)          # fills en_min/en_max/v_min/v_max/deltas
        check(lib, h[], ccall((:ptmc_reset_counters, lib), Cint, (Ptr{Cvoid},), h[]))
        check(lib, h[], ccall((:ptmc_set_hist_bounds, lib), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Cint),
                              h[], results.en_min, results.en_max, results.v_min, results.v_max, 0))
        # main loop (src/MCRun.jl:397-421) in chunks that end where the reference's driver would look at the states
        stops = filter(x -> x !== false, (save, saveconfigs, rdfsave ? 1000 : false))
        rdfsave && check(lib, h[], ccall((:ptmc_rdf_enable, lib), Cint, (Ptr{Cvoid}, Float64), h[], results.delta_r2))
        i = 1
        while i <= mc_params.mc_cycles
            k = min(b.chunk, mc_params.mc_cycles - i + 1)
            for s in stops; k = min(k, s - rem(i - 1, s)); end
            check(lib, h[], ccall((:ptmc_run_cycles, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Cint), h[], i, k, 1))
            i += k
            # update_rdf! (src/Sampling.jl:343-347): sampled cycles with rem(idx, 1000) == 0
            rdfsave && rem(i - 1, 1000) == 0 && rem(i - 1, mc_params.mc_sample) == 0 &&
                check(lib, h[], ccall((:ptmc_rdf_sample, lib), Cint, (Ptr{Cvoid},), h[]))
            if any(s -> s !== 1000 && rem(i - 1, s) == 0, filter(x -> x !== false, (save, saveconfigs)))
                download!(lib, h[], mc_states, results, N, T, L, npt)
                rdfsave && download_rdf!(lib, h[], results, T, L)
                save !== false && rem(i - 1, save) == 0 && ParallelTemperingMonteCarlo.checkpoint(i - 1, mc_states, results, ensemble, rdfsave)
                saveconfigs !== false && rem(i - 1, saveconfigs) == 0 && ParallelTemperingMonteCarlo.save_configs(mc_states, string(configsname, i - 1))
            end
        end
        download!(lib, h[], mc_states, results, N, T, L, npt)
        rdfsave && download_rdf!(lib, h[], results, T, L)
    finally
        ccall((:ptmc_destroy, lib), Cvoid, (Ptr{Cvoid},), h[])
    end
    results = finalise_results(mc_states, mc_params, results)
    return mc_states, results
end

# state out: refill the host MCStates of ladder 1 and the histograms of `results`
function download!(lib, h, mc_states, results, N, T, L, npt)
    pos = Array{Float64}(undef, 3, N, T, L); box = Array{Float64}(undef, 2, T, L); en = Array{Float64}(undef, T, L)
    md = Array{Float64}(undef, 4, T, L); cnt = Array{Int64}(undef, PTMC_NCOUNT, T, L); ham = Array{Float64}(undef, 2, T, L)
    check(lib, h, ccall((:ptmc_download_states, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                        h, pos, box, en, md, cnt, ham))
    for t in 1:T
        s = mc_states[t]
        for j in 1:N; s.config.positions[j] = typeof(s.config.positions[j])(pos[1, j, t, 1], pos[2, j, t, 1], pos[3, j, t, 1]); end
        s.en_tot = en[t, 1]; s.max_displ .= md[:, t, 1]; s.ham .= ham[:, t, 1]
        s.count_atom .= cnt[1:2, t, 1]; s.count_vol .= cnt[3:4, t, 1]; s.count_exc .= cnt[9:10, t, 1]
        ParallelTemperingMonteCarlo.get_distance2_mat!(s.dist2_mat, s.config)     # derived state, recomputed on the host
    end
    nb2 = results.n_bin + 2
    if npt
        ev = Array{Float64}(undef, nb2, nb2, T, L)       # row-major [L][T][E][V] -> Julia (V, E, T, L)
        check(lib, h, ccall((:ptmc_download_hist, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h, C_NULL, ev, 0))
        for t in 1:T; results.ev_histogram[t] .= permutedims(ev[:, :, t, 1]); end
    else
        eh = Array{Float64}(undef, nb2, T, L)
        check(lib, h, ccall((:ptmc_download_hist, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h, eh, C_NULL, 0))
        for t in 1:T; results.en_histogram[t] .= eh[:, t, 1]; end
    end
    return nothing
end

function download_rdf!(lib, h, results, T, L)
    rdf = Array{Float64}(undef, 5 * results.n_bin, T, L)
    check(lib, h, ccall((:ptmc_download_rdf, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, rdf))
    for t in 1:T; results.rdf[t] .= rdf[:, t, 1]; end
    return nothing
end

# RuNNer forward pass: point the package's `lib_path()` at the directory holding a copy/symlink of libptmc_b200.so named
# librunnerjulia.so; the `forward` symbol has the Fortran library's exact ABI (MachineLearningPotential/ForwardPass.jl:60-83).

end # module
