# PTMCB200.jl — the binding a maintainer of ParallelTemperingMonteCarlo.jl adds to run the sweep hot path on a B200
# through libptmc_b200.so (C ABI: include/ptmc_b200.h).  Thin by design: it converts the package's own types into the
# POD config + flat arrays of the ABI, calls `ccall`, and fills the package's own `MCState`s / `Output` back.
#
# NOT executed in the build image (no Julia there).  The same call sequence is exercised from Python by
# paralleltemperingmontecarlo.jl_b200/engine.py + host.py (ptmc_run) and from plain C by tests/c/npt_sequence.c
# (built with gcc against the header, run on the GPU by tests/test_c_sequence.py).
#
#   using ParallelTemperingMonteCarlo, PTMCB200
#   backend = B200Backend(lib = "/path/to/libptmc_b200.so", device = 0)
#   states, results = ptmc_run!(backend, mc_params, temp, start_config, potential, ensemble; save = 1000)
#   # many independent ladders (seeds) in one batch:
#   backend = B200Backend(lib = ..., n_ladders = 4096, shared_bounds = true)
#   all_states, all_results, summed = ptmc_run_batch!(backend, mc_params, temp, start_config, potential, ensemble)
module PTMCB200

using ParallelTemperingMonteCarlo
using ParallelTemperingMonteCarlo: MCParams, TempGrid, Config, MCState, Output, NVT, NPT, NNVT,
    SphericalBC, CubicBC, RhombicBC, RectangularBC, ELJPotential, ELJPotentialEven, ELJPotentialB, LookupTablePotential,
    EmbeddedAtomPotential, RuNNerPotential, RuNNerPotential2Atom,
    initialisation, finalise_results, initialise_histograms!, check_e_bounds, get_volume, get_r_cut,
    save_init, save_histparams, checkpoint, save_configs

export B200Backend, ptmc_run!, ptmc_run_batch!

const PTMC_ABI_VERSION = Int32(4)
const PTMC_MAX_COEFF = 16
const PTMC_NCOUNT = 10

Base.@kwdef struct B200Backend
    lib::String
    device::Int = 0
    n_ladders::Int = 1          # independent ladders (seeds) batched on the device
    seed::UInt64 = 1234
    rng_mode::Int32 = 2         # PTMC_RNG_BATCH8
    lanes_per_replica::Int32 = 0
    chunk::Int = 100_000        # cycles per ccall (the driver touches states only at save / saveconfigs cadence)
    shared_bounds::Bool = false # all ladders bin into one [en_min, en_max] (min / max over every ladder): histograms can be summed
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
    n_atom_swaps::Int32        # ABI 4: NNVT.n_atom_swaps, RuNNerPotential2Atom descriptor
    runner2::Ptr{Cvoid}
end

bc_kind(::SphericalBC) = Int32(0); bc_kind(::CubicBC) = Int32(1); bc_kind(::RhombicBC) = Int32(2); bc_kind(::RectangularBC) = Int32(3)
box_of(bc::SphericalBC) = (bc.radius2, 0.0)
box_of(bc::CubicBC) = (bc.box_length, bc.box_length)
box_of(bc::Union{RhombicBC,RectangularBC}) = (bc.box_length, bc.box_height)
# the boundary condition a slot holds after volume moves: rebuilt from the downloaded box {box_length, box_height}
rebuild_bc(bc::SphericalBC, L, h) = bc
rebuild_bc(::CubicBC, L, h) = CubicBC(L)
rebuild_bc(::RhombicBC, L, h) = RhombicBC(L, h)
rebuild_bc(::RectangularBC, L, h) = RectangularBC(L, h)

ens_kind(::NVT) = Int32(0); ens_kind(::NPT) = Int32(1); ens_kind(::NNVT) = Int32(2)

pot_fields(p::ELJPotential{N}) where {N} = (Int32(0), Int32(N), Tuple(p.coeff), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::ELJPotentialEven{N}) where {N} = (Int32(1), Int32(N), Tuple(p.coeff), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::EmbeddedAtomPotential) = (Int32(2), Int32(0), (), (p.n, p.m, p.ean, p.eCam))
pot_fields(p::LookupTablePotential) = (Int32(5), Int32(0), (), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::ELJPotentialB{N}) where {N} = (Int32(4), Int32(N), Tuple(p.coeff_c), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::RuNNerPotential) = (Int32(3), Int32(0), (), (0.0, 0.0, 0.0, 0.0))
pot_fields(p::RuNNerPotential2Atom) = (Int32(6), Int32(0), (), (0.0, 0.0, 0.0, 0.0))

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

# mirror of `ptmc_runner2_desc`
struct PtmcRunner2Desc
    n_rad::Int32
    n_ang::Int32
    n1::Int32
    n2::Int32
    rad_eta::Ptr{Float64}
    ang_eta::Ptr{Float64}
    ang_lambda::Ptr{Float64}
    ang_zeta::Ptr{Float64}
    gmax_gmin_1::Ptr{Float64}
    gmax_gmin_2::Ptr{Float64}
    r_cut::Float64
    boundary2::Float64
    n_layers::Int32
    num_nodes::Ptr{Int32}
    activation::Ptr{Int32}
    n_params::Int32
    params_1::Ptr{Float64}
    params_2::Ptr{Float64}
end

# The package's symmetry-function structs hold G_norm = 1 / (G_max - G_min) and G_offset = -G_min * G_norm
# (SymmFunc.jl:40-44,107-112), the descriptors want the scaling.data pair itself: a script that still has its `G_value_vec`
# should pass that instead (`scaling = G_value_vec`, for two species `scaling = (G_value_vec, G_value_vec_b)`), because the pair
# recovered here, G_min = -G_offset / G_norm and G_max = G_min + 1 / G_norm, can differ from the file's by an ulp.
gpair(norm, off) = (gmin = -off / norm; (gmin + 1 / norm, gmin))

# RuNNerPotential -> descriptor + the arrays it points into (keep them alive across ptmc_create)
function runner_desc(p::RuNNerPotential{Nrad,Nang}; scaling=nothing) where {Nrad,Nang}
    rs, as = p.radsymfunctions, p.angsymfunctions
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

# RuNNerPotential2Atom -> descriptor: rows in g-matrix order (radial k: rows 2k-1, 2k; angular k: 2 Nrad + 3(k-1) + 1..3),
# species 1 = entries 1:2 / 1:3 of G_norm / tpz, species 2 = entries 3:4 / 4:6 (SymmFunc.jl:64-83,133-171)
function runner2_desc(p::RuNNerPotential2Atom{Nrad,Nang,N1,N2}; scaling=nothing) where {Nrad,Nang,N1,N2}
    rs, as = p.radsymfunctions, p.angsymfunctions
    g1 = Float64[]; g2 = Float64[]
    if scaling === nothing
        for k in 1:Nrad, q in 1:2
            append!(g1, gpair(rs.G_norm[k][q], rs.G_offset[k][q])); append!(g2, gpair(rs.G_norm[k][2 + q], rs.G_offset[k][2 + q]))
        end
        for k in 1:Nang, q in 1:3
            f = 2.0^(1 - as.zeta[k])
            append!(g1, gpair(as.tpz[k][q] / f, as.G_offset[k][q])); append!(g2, gpair(as.tpz[k][3 + q] / f, as.G_offset[k][3 + q]))
        end
    else
        for row in 1:(2 * Nrad + 3 * Nang)
            append!(g1, (scaling[1][row][1], scaling[1][row][2])); append!(g2, (scaling[2][row][1], scaling[2][row][2]))
        end
    end
    arrays = (collect(Float64, rs.eta), collect(Float64, as.eta), collect(Float64, as.lambda), collect(Float64, as.zeta), g1, g2,
              collect(Int32, p.nnp1.num_nodes), collect(Int32, p.nnp1.activation_functions),
              collect(Float64, p.nnp1.parameters), collect(Float64, p.nnp2.parameters))
    desc = PtmcRunner2Desc(Nrad, Nang, N1, N2, pointer(arrays[1]), pointer(arrays[2]), pointer(arrays[3]), pointer(arrays[4]),
                           pointer(arrays[5]), pointer(arrays[6]), p.r_cut, p.boundary, p.nnp1.n_layers, pointer(arrays[7]),
                           pointer(arrays[8]), p.nnp1.n_params, pointer(arrays[9]), pointer(arrays[10]))
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
`(mc_states, results)` — those of the first ladder.  The sweep (`mc_cycle!`, src/MCRun.jl:192-253) runs on the device in
chunks; states come back to the host only where the reference's driver touches them (histogram bounds after
equilibration, checkpoints, end of run).
"""
function ParallelTemperingMonteCarlo.ptmc_run!(b::B200Backend, mc_params::MCParams, temp::TempGrid, start_config,
                                               potential, ensemble; kwargs...)
    all_states, all_results, _ = ptmc_run_batch!(b, mc_params, temp, start_config, potential, ensemble; kwargs...)
    return all_states[1], all_results[1]
end

"""
    ptmc_run_batch!(backend, mc_params, temp, start_config, potential, ensemble; ...) -> (states, results, summed)

`backend.n_ladders` independent ladders in one batch: `states[l]` / `results[l]` are the `MCState` vector and the `Output` of
ladder `l` (each with its own histogram bounds unless `backend.shared_bounds`), `summed` is an `Output` whose histograms are the
sum over all ladders (reduced on the device; meaningful with `shared_bounds = true`).  Checkpoints of ladder `l` go to
`./checkpoint` for a single ladder and to `./ladder<l>/checkpoint` otherwise.
"""
function ptmc_run_batch!(b::B200Backend, mc_params::MCParams, temp::TempGrid, start_config, potential, ensemble;
                         rdfsave=false, save=false, saveconfigs=false, configsname="configuration", workingdirectory=pwd(), kwargs...)
    lib = b.lib
    cd(workingdirectory)
    T, N, L = mc_params.n_traj, mc_params.n_atoms, b.n_ladders
    ladder_dir(l) = L == 1 ? "." : "ladder$l"
    in_ladder_dir(f, l) = (mkpath(ladder_dir(l)); cd(f, ladder_dir(l)))
    if save !== false
        for l in 1:L; in_ladder_dir(() -> save_init(potential, ensemble, mc_params, temp), l); end      # src/MCRun.jl:378-380
    end
    # the package's own initialisation builds MCStates (host copies), MoveStrategy and the empty Output - once per ladder
    inits = [initialisation(mc_params, temp, start_config, potential, ensemble) for _ in 1:L]
    all_states = [x[1] for x in inits]
    all_results = [x[3] for x in inits]
    bc = all_states[1][1].config.boundary_condition
    pk, nc, coeff, eam = pot_fields(potential)
    npt = ensemble isa NPT
    betas = collect(Float64, temp.beta_grid)
    ea = potential isa ELJPotentialB ? collect(Float64, potential.coeff_a) : Float64[]
    eb = potential isa ELJPotentialB ? collect(Float64, potential.coeff_b) : Float64[]
    islut = potential isa LookupTablePotential
    lut = islut ? collect(Float64, permutedims(potential.table)) : Float64[]        # memory order: distance index fastest
    scaling = get(kwargs, :scaling, nothing)
    rdesc, rkeep = potential isa RuNNerPotential ? runner_desc(potential; scaling=scaling) : (nothing, ())
    r2desc, r2keep = potential isa RuNNerPotential2Atom ? runner2_desc(potential; scaling=scaling) : (nothing, ())
    cfg = Ref(PtmcConfig(PTMC_ABI_VERSION, b.device, L, T, N, ens_kind(ensemble), ensemble.n_atom_moves,
        npt ? ensemble.n_volume_moves : 0, bc_kind(bc), npt ? ensemble.pressure : 0.0, pk, nc, pad16(coeff),
        eam..., rdesc === nothing ? Ptr{Cvoid}(C_NULL) : Ptr{Cvoid}(Base.unsafe_convert(Ptr{PtmcRunnerDesc}, rdesc)), pointer(betas), mc_params.n_adjust, mc_params.mc_sample, mc_params.n_bin, b.rng_mode,
        mc_params.min_acc, mc_params.max_acc, b.seed, b.lanes_per_replica, (npt && ensemble.separated_volume) ? 1 : 0,
        isempty(ea) ? Ptr{Float64}(C_NULL) : pointer(ea), isempty(eb) ? Ptr{Float64}(C_NULL) : pointer(eb),
        islut ? pointer(lut) : Ptr{Float64}(C_NULL), islut ? potential.l_dist : Int32(0), islut ? potential.l_angle : Int32(0),
        islut ? potential.start_dist : 0.0, islut ? potential.d_dist : 0.0, islut ? potential.start_angle : 0.0,
        islut ? potential.d_angle : 0.0, islut ? potential.c6coeff : 0.0,
        ensemble isa NNVT ? ensemble.n_atom_swaps : 0,
        r2desc === nothing ? Ptr{Cvoid}(C_NULL) : Ptr{Cvoid}(Base.unsafe_convert(Ptr{PtmcRunner2Desc}, r2desc))))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve betas ea eb lut rdesc rkeep r2desc r2keep begin
        rc = ccall((:ptmc_create, lib), Cint, (Ref{PtmcConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    end
    rc == 0 || error("libptmc_b200: " * unsafe_string(ccall((:ptmc_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    summed = Output{Float64}(mc_params.n_bin)
    try
        # state in: positions [L][T][N][3] (row-major = Julia (3, N, T, L)), box [L][T][2]
        pos = Array{Float64}(undef, 3, N, T, L); box = Array{Float64}(undef, 2, T, L)
        for l in 1:L, t in 1:T
            for (j, p) in enumerate(all_states[l][t].config.positions); pos[:, j, t, l] .= p; end
            box[:, t, l] .= box_of(all_states[l][t].config.boundary_condition)
        end
        check(lib, h[], ccall((:ptmc_upload_states, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h[], pos, box, C_NULL))
        # equilibration_cycle! (src/MCRun.jl:282-313): MC for eq_cycles ÷ 2 cycles, then the energy bounds of every ladder
        n_eq = mc_params.eq_cycles ÷ 2
        n_eq > 0 && check(lib, h[], ccall((:ptmc_run_cycles, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Cint), h[], 1, n_eq, 0))
        en = Array{Float64}(undef, T, L)
        check(lib, h[], ccall((:ptmc_download_states, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                              h[], C_NULL, C_NULL, en, C_NULL, C_NULL, C_NULL))
        bounds = Vector{Vector{Float64}}(undef, L)
        for l in 1:L
            ebounds = [100.0, -100.0]
            for _ in (n_eq + 1):mc_params.eq_cycles, t in 1:T
                ebounds = check_e_bounds(en[t, l], ebounds)
            end
            bounds[l] = ebounds
        end
        if b.shared_bounds
            lo, hi = minimum(x -> x[1], bounds), maximum(x -> x[2], bounds)
            for l in 1:L; bounds[l] = [lo, hi]; end
        end
        for l in 1:L
            initialise_histograms!(mc_params, all_results[l], bounds[l], bc)     # fills en_min/en_max/v_min/v_max/deltas
        end
        initialise_histograms!(mc_params, summed, [minimum(x -> x[1], bounds), maximum(x -> x[2], bounds)], bc)
        check(lib, h[], ccall((:ptmc_reset_counters, lib), Cint, (Ptr{Cvoid},), h[]))
        emin = [r.en_min for r in all_results]; emax = [r.en_max for r in all_results]
        vmin = [r.v_min for r in all_results]; vmax = [r.v_max for r in all_results]
        check(lib, h[], ccall((:ptmc_set_hist_bounds, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                              h[], emin, emax, vmin, vmax, 1))                   # per_ladder = 1
        if save !== false
            for l in 1:L; in_ladder_dir(() -> save_histparams(all_results[l]), l); end                  # src/MCRun.jl:390-392
        end
        # main loop (src/MCRun.jl:397-421) in chunks that end where the reference's driver would look at the states
        stops = filter(x -> x !== false, (save, saveconfigs, rdfsave ? 1000 : false))
        rdfsave && check(lib, h[], ccall((:ptmc_rdf_enable, lib), Cint, (Ptr{Cvoid}, Float64), h[], all_results[1].delta_r2))
        function write_out(idx, do_save, do_configs)
            download!(lib, h[], all_states, all_results, N, T, L, npt)
            rdfsave && download_rdf!(lib, h[], all_results, T, L)
            for l in 1:L
                in_ladder_dir(l) do
                    do_save && checkpoint(idx, all_states[l], all_results[l], ensemble, rdfsave)
                    do_configs && save_configs(all_states[l], string(configsname, idx))
                end
            end
        end
        i = 1
        last_saved = 0
        while i <= mc_params.mc_cycles
            k = min(b.chunk, mc_params.mc_cycles - i + 1)
            for s in stops; k = min(k, s - rem(i - 1, s)); end
            check(lib, h[], ccall((:ptmc_run_cycles, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Cint), h[], i, k, 1))
            i += k
            done = i - 1
            # update_rdf! (src/Sampling.jl:343-347): sampled cycles with rem(idx, 1000) == 0
            rdfsave && rem(done, 1000) == 0 && rem(done, mc_params.mc_sample) == 0 &&
                check(lib, h[], ccall((:ptmc_rdf_sample, lib), Cint, (Ptr{Cvoid},), h[]))
            do_save = save !== false && rem(done, save) == 0                      # src/MCRun.jl:409-411
            do_configs = saveconfigs !== false && rem(done, saveconfigs) == 0     # :412-414
            if do_save || do_configs
                write_out(done, do_save, do_configs)
                do_save && (last_saved = done)
            end
        end
        if save !== false && last_saved != mc_params.mc_cycles                     # :424-427
            write_out(mc_params.mc_cycles, true, false)
        end
        download!(lib, h[], all_states, all_results, N, T, L, npt)
        rdfsave && download_rdf!(lib, h[], all_results, T, L)
        download_summed!(lib, h[], summed, T, npt)
    finally
        ccall((:ptmc_destroy, lib), Cvoid, (Ptr{Cvoid},), h[])
    end
    for l in 1:L
        all_results[l] = finalise_results(all_states[l], mc_params, all_results[l])
    end
    return all_states, all_results, summed
end

# state out: refill the host MCStates and the histograms of every ladder
function download!(lib, h, all_states, all_results, N, T, L, npt)
    pos = Array{Float64}(undef, 3, N, T, L); box = Array{Float64}(undef, 2, T, L); en = Array{Float64}(undef, T, L)
    md = Array{Float64}(undef, 4, T, L); cnt = Array{Int64}(undef, PTMC_NCOUNT, T, L); ham = Array{Float64}(undef, 2, T, L)
    check(lib, h, ccall((:ptmc_download_states, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                        h, pos, box, en, md, cnt, ham))
    nb2 = all_results[1].n_bin + 2
    ev = npt ? Array{Float64}(undef, nb2, nb2, T, L) : nothing       # row-major [L][T][E][V] -> Julia (V, E, T, L)
    eh = npt ? nothing : Array{Float64}(undef, nb2, T, L)
    if npt
        check(lib, h, ccall((:ptmc_download_hist, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h, C_NULL, ev, 0))
    else
        check(lib, h, ccall((:ptmc_download_hist, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h, eh, C_NULL, 0))
    end
    for l in 1:L, t in 1:T
        s = all_states[l][t]
        # Config is immutable: a new one with the positions AND the boundary condition the slot holds now (volume moves change
        # the box; checkpoint_config prints bc_info(state.config.boundary_condition), src/ReadSave.jl:142-158)
        newbc = rebuild_bc(s.config.boundary_condition, box[1, t, l], box[2, t, l])
        s.config = Config([typeof(s.config.positions[j])(pos[1, j, t, l], pos[2, j, t, l], pos[3, j, t, l]) for j in 1:N], newbc)
        if npt
            s.ensemble_variables.r_cut = get_r_cut(newbc)
        end
        s.en_tot = en[t, l]; s.max_displ .= md[:, t, l]; s.ham .= ham[:, t, l]
        s.count_atom .= cnt[1:2, t, l]; s.count_vol .= cnt[3:4, t, l]
        s.count_vol_xy .= cnt[5:6, t, l]; s.count_vol_z .= cnt[7:8, t, l]       # separated volume moves
        s.count_exc .= cnt[9:10, t, l]
        ParallelTemperingMonteCarlo.get_distance2_mat!(s.dist2_mat, s.config)     # derived state, recomputed on the host
        if npt
            all_results[l].ev_histogram[t] .= permutedims(ev[:, :, t, l])
        else
            all_results[l].en_histogram[t] .= eh[:, t, l]
        end
    end
    return nothing
end

# histograms summed over the ladder axis on the device (ptmc_download_hist(..., reduce_ladders = 1))
function download_summed!(lib, h, summed, T, npt)
    nb2 = summed.n_bin + 2
    if npt
        ev = Array{Float64}(undef, nb2, nb2, T)
        check(lib, h, ccall((:ptmc_download_hist, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h, C_NULL, ev, 1))
        for t in 1:T; summed.ev_histogram[t] .= permutedims(ev[:, :, t]); end
    else
        eh = Array{Float64}(undef, nb2, T)
        check(lib, h, ccall((:ptmc_download_hist, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h, eh, C_NULL, 1))
        for t in 1:T; summed.en_histogram[t] .= eh[:, t]; end
    end
    return nothing
end

function download_rdf!(lib, h, all_results, T, L)
    rdf = Array{Float64}(undef, 5 * all_results[1].n_bin, T, L)
    check(lib, h, ccall((:ptmc_download_rdf, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, rdf))
    for l in 1:L, t in 1:T; all_results[l].rdf[t] .= rdf[:, t, l]; end
    return nothing
end

# RuNNer forward pass: point the package's `lib_path()` at the directory holding a copy/symlink of libptmc_b200.so named
# librunnerjulia.so; the `forward` symbol has the Fortran library's exact ABI (MachineLearningPotential/ForwardPass.jl:60-83).

end # module
