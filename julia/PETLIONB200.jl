# PETLIONB200.jl -- the `ccall` layer a PETLION.jl maintainer would add (e.g. as src/b200.jl) to run the hot
# path -- initialize_simulation! + IDA + solve! + exit_simulation! (src/model_evaluation.jl:174-382) and the
# residual/Jacobian callbacks (src/physics_equations/scalar_residual.jl:558-602) -- on B200 GPUs through
# libpetlion_b200.so (C ABI: include/petlion_b200.h).
#
# Julia is not part of the build image, so this file is UNTESTED here (it is kept parseable and in step with the
# header by tests/test_host_cpu.py::test_julia_shim_covers_the_header); the tested stand-in with the same structure is
# petlion.jl_b200/api.py.  Arrays are batch-major: a Julia Matrix{Float64} of size (N, B) -- one COLUMN per system --
# has exactly the memory layout the library expects ([B][N], one contiguous row per system).
module PETLIONB200

const lib = "libpetlion_b200"          # petlion.jl_b200/csrc/libpetlion_b200.so on the library path

struct ModelDesc
    cathode::Cint; N_p::Cint; N_s::Cint; N_n::Cint; N_a::Cint; N_z::Cint; N_r_p::Cint; N_r_n::Cint
    temperature::Cint; aging::Cint; device::Cint
    rxn_p::Cint; rxn_n::Cint          # 0 rxn_BV, 1 rxn_MHC
    fickian_spectral::Cint            # p.numerics.Fickian_method === :spectral
end
struct Run
    method::Cint; input_kind::Cint; value::Cdouble; tf::Cdouble; new_run::Cint; reserved::Cint
end
struct Opts
    abstol::Cdouble; reltol::Cdouble; abstol_init::Cdouble; reltol_init::Cdouble
    maxiters::Cint; check_bounds::Cint; interp_final::Cint; skip_alg_deriv::Cint
end
struct Bounds
    V_max::Cdouble; V_min::Cdouble; SOC_max::Cdouble; SOC_min::Cdouble; T_max::Cdouble; c_s_n_max::Cdouble
    I_max::Cdouble; I_min::Cdouble; η_plating_min::Cdouble; c_e_min::Cdouble; dfilm_max::Cdouble
end
struct Summary
    t_end::Cdouble; V_end::Cdouble; I_end::Cdouble; SOC_end::Cdouble; T_end::Cdouble; aux_end::Cdouble
    flag::Cint; n_steps::Cint; n_res::Cint; n_jac::Cint; n_netf::Cint; n_ncfn::Cint; n_newton_init::Cint; n_reinit::Cint
end
struct InputTable
    n::Cint; t::Ptr{Cdouble}; v::Ptr{Cdouble}; n_tdiscon::Cint; tdiscon::Ptr{Cdouble}
end

const METHOD = Dict(:I => 0, :V => 1, :P => 2, :dT => 3, :η_p => 4,       # scalar_residual.jl:167-202
                    :dc_s_p_max => 6, :dc_s_p_min => 7, :dc_s_n_max => 8, :dc_s_n_min => 9,   # input_methods.jl:190-245
                    :dc_e_max => 10, :dc_e_min => 11)                       # (simulate! on isothermal models without aging)
const INPUT_VALUE, INPUT_HOLD, INPUT_REST = 0, 1, 2                      # input_methods.jl:5-74
const MEM_HOST, MEM_DEVICE = 0, 1

check(rc) = rc == 0 || error(unsafe_string(ccall((:plb_last_error, lib), Cstring, ())))

# ---- model handle: petlion(...) (src/external.jl:2-18) ------------------------------------------------------------
rxn_code(f) = Symbol(f) === :rxn_MHC ? 1 : 0
cathode_code(c::Symbol) = c === :LCO ? 0 : (c === :NMC_LGM50 ? 2 : 1)          # PLB_CATHODE_*
function create(cathode::Symbol, N, temperature::Bool, aging; device::Integer = 0, rxn_p = :rxn_BV, rxn_n = :rxn_BV, Fickian_method = :finite_difference)::Ptr{Cvoid}
    d = ModelDesc(cathode_code(cathode), N.p, N.s, N.n, N.a, N.z, N.r_p, N.r_n, temperature, aging === :SEI, device,
                  rxn_code(rxn_p), rxn_code(rxn_n), Fickian_method === :spectral)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:plb_create, lib), Cint, (Ref{ModelDesc}, Ref{Ptr{Cvoid}}), d, h))
    return h[]
end
destroy(h) = ccall((:plb_destroy, lib), Cint, (Ptr{Cvoid},), h)
set_stream(h, stream::Ptr{Cvoid}) = check(ccall((:plb_set_stream, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h, stream))
nstates(h) = ccall((:plb_nstates, lib), Cint, (Ptr{Cvoid},), h)
ndiff(h) = ccall((:plb_ndiff, lib), Cint, (Ptr{Cvoid},), h)
ntheta(h) = ccall((:plb_ntheta, lib), Cint, (Ptr{Cvoid},), h)
jac_nnz(h, method) = ccall((:plb_jac_nnz, lib), Cint, (Ptr{Cvoid}, Cint), h, METHOD[method])
launch_count(h) = ccall((:plb_launch_count, lib), Clonglong, (Ptr{Cvoid},), h)
last_kernel_ms(h) = ccall((:plb_last_kernel_ms, lib), Cfloat, (Ptr{Cvoid},), h)
function variant_info(family::Integer)
    out = zeros(Clonglong, 8)
    ccall((:plb_variant_info, lib), Cint, (Cint, Ptr{Clonglong}), family, out)
    return out
end

# θ_keys of the generated functions (generate_functions.jl:327-363, 387) and the defaults of src/params.jl
function theta_keys(h)
    keys = Vector{Cstring}(undef, ntheta(h))
    ccall((:plb_theta_keys, lib), Cint, (Ptr{Cvoid}, Ptr{Cstring}), h, keys)
    return Symbol.(unsafe_string.(keys))
end
theta_index(h, key::Symbol) = ccall((:plb_theta_index, lib), Cint, (Ptr{Cvoid}, Cstring), h, String(key)) + 1
function theta_defaults(h)
    row = zeros(ntheta(h))
    check(ccall((:plb_theta_defaults, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h, row))
    return row
end
function bounds_defaults(h)
    b = Ref{Bounds}()
    check(ccall((:plb_bounds_defaults, lib), Cint, (Ptr{Cvoid}, Ref{Bounds}), h, b)); b[]
end
function opts_defaults(h)
    o = Ref{Opts}()
    check(ccall((:plb_opts_defaults, lib), Cint, (Ptr{Cvoid}, Ref{Opts}), h, o)); o[]
end
function calc_I1C(h, θ::Matrix{Float64})           # auxiliary_states_and_coefficients.jl:631-647, one value per column
    out = zeros(size(θ, 2))
    check(ccall((:plb_calc_I1C, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), h, size(θ, 2), θ, out)); out
end

# J_full.sp (scalar_residual.jl:501): CSC pattern, 1-based like SparseMatrixCSC
function jac_pattern(h, method::Symbol)
    colptr = zeros(Cint, nstates(h) + 1); rowval = zeros(Cint, jac_nnz(h, method))
    check(ccall((:plb_jac_pattern, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Cint), h, METHOD[method], colptr, rowval, 1))
    return colptr, rowval
end

# ---- operator level: the callback surface over a batch ---------------------------------------------------------------
# initial_guess!(Y0, SOC, θ_tot, X_applied) -- model_evaluation.jl:204
initial_guess!(h, B, soc, θ, Y0; mem = MEM_HOST) =
    check(ccall((:plb_initial_guess, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint), h, B, soc, θ, Y0, mem))

# R_full / J_full (scalar_residual.jl:558-602); res or nzval may be C_NULL
resjac!(h, B, Y, YP, γ, θ, run::Run, values, res, nzval; mem = MEM_HOST) =
    check(ccall((:plb_resjac, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Run}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
        h, B, Y, YP, γ, θ, run, values, res, nzval, mem))

# newtons_method! (model_evaluation.jl:430-480)
newton_init!(h, B, Y, YP, θ, run::Run, values, opts::Opts, status; mem = MEM_HOST) =
    check(ccall((:plb_newton_init, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Run}, Ptr{Cdouble}, Ref{Opts}, Ptr{Cint}, Cint),
        h, B, Y, YP, θ, run, values, opts, status, mem))

# KLU's role (factorize! + \, model_evaluation.jl:417-428): x = (∂F/∂Y + γ ∂F/∂Y′)⁻¹ rhs
linear_solve!(h, B, Y, YP, γ, θ, run::Run, values, rhs, x, status; mem = MEM_HOST) =
    check(ccall((:plb_linear_solve, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Run}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
        h, B, Y, YP, γ, θ, run, values, rhs, x, status, mem))

# ---- integrator level: simulate / simulate! -----------------------------------------------------------------------
# p.opts.tstops (params.jl:272) and the results at user times of simulate(p, tf::AbstractVector) (model_evaluation.jl:80):
# both apply to the following simulate call
set_tstops(h, tstops::Vector{Float64}) =
    check(ccall((:plb_set_tstops, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h, length(tstops), tstops))
set_dense_output(h, t::Vector{Float64}, V, I, SOC, T, Y, n_done; mem = MEM_HOST) =
    check(ccall((:plb_set_dense_output, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
        h, length(t), t, V, I, SOC, T, Y, n_done, mem))

# replaces initialize_simulation! + solve! + exit_simulation!; run.new_run = 0 is simulate!(sol, p, ...): the state
# arrays Y, YP, SOC, t are the hand-off.  trY (or C_NULL): every saved row's full state (outputs = :all)
simulate!(h, B, θ, run::Run, values, opts::Opts, bounds::Bounds, soc0, Y, YP, SOC, t, summ, nsave, trt, trV, trI, trS, trT, trY, trn; mem = MEM_HOST) =
    check(ccall((:plb_simulate, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Run}, Ptr{Cdouble}, Ref{Opts}, Ref{Bounds}, Ptr{Cdouble},
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Summary}, Cint,
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
        h, B, θ, run, values, opts, bounds, soc0, Y, YP, SOC, t, summ, nsave, trt, trV, trI, trS, trT, trY, trn, mem))

# run_function inputs (structures.jl:55-63): the closure is tabulated on the host -- knots where it has kinks, a repeated
# knot time where it jumps -- and `scale` (or C_NULL) multiplies it per system
function simulate_table!(h, B, θ, run::Run, tt::Vector{Float64}, vv::Vector{Float64}, tdiscon::Vector{Float64}, scale,
                         opts::Opts, bounds::Bounds, soc0, Y, YP, SOC, t, summ, nsave, trt, trV, trI, trS, trT, trY, trn; mem = MEM_HOST)
    GC.@preserve tt vv tdiscon begin
        tab = InputTable(length(tt), pointer(tt), pointer(vv), length(tdiscon), pointer(tdiscon))
        check(ccall((:plb_simulate_table, lib), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Run}, Ref{InputTable}, Ptr{Cdouble}, Ref{Opts}, Ref{Bounds}, Ptr{Cdouble},
             Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Summary}, Cint,
             Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}, Cint),
            h, B, θ, run, tab, scale, opts, bounds, soc0, Y, YP, SOC, t, summ, nsave, trt, trV, trI, trS, trT, trY, trn, mem))
    end
end

# ---- several GPUs of the box behind one call (contiguous batch shards, one ncclAllGather of the summaries) --------
function group_create(cathode::Symbol, N, temperature::Bool, aging, devices::Vector{<:Integer}; rxn_p = :rxn_BV, rxn_n = :rxn_BV, Fickian_method = :finite_difference)::Ptr{Cvoid}
    d = ModelDesc(cathode_code(cathode), N.p, N.s, N.n, N.a, N.z, N.r_p, N.r_n, temperature, aging === :SEI, 0,
                  rxn_code(rxn_p), rxn_code(rxn_n), Fickian_method === :spectral)
    g = Ref{Ptr{Cvoid}}()
    check(ccall((:plb_group_create, lib), Cint, (Ref{ModelDesc}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), d, length(devices), Cint.(devices), g))
    return g[]
end
group_destroy(g) = ccall((:plb_group_destroy, lib), Cint, (Ptr{Cvoid},), g)
group_size(g) = ccall((:plb_group_size, lib), Cint, (Ptr{Cvoid},), g)
group_handle(g, k::Integer) = ccall((:plb_group_handle, lib), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), g, k)
group_simulate!(g, B, θ, run::Run, values, opts::Opts, bounds::Bounds, soc0, Y, YP, SOC, t, summ, nsave, trt, trV, trI, trS, trT, trn) =
    check(ccall((:plb_group_simulate, lib), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ref{Run}, Ptr{Cdouble}, Ref{Opts}, Ref{Bounds}, Ptr{Cdouble},
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Summary}, Cint,
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
        g, B, θ, run, values, opts, bounds, soc0, Y, YP, SOC, t, summ, nsave, trt, trV, trI, trS, trT, trn))
function group_device_summaries(g, k::Integer)       # device pointer on device k to the gathered summaries, rows per device
    out = Ref{Ptr{Summary}}(); rows = Ref{Cint}()
    check(ccall((:plb_group_device_summaries, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Summary}}, Ref{Cint}), g, k, out, rows))
    return out[], rows[]
end
group_last_gather_ms(g) = ccall((:plb_group_last_gather_ms, lib), Cfloat, (Ptr{Cvoid},), g)

# ---- the method PETLION.simulate would gain: a batch of parameter columns in, one solution per column out ----------
"""
    simulate_batch(h, θ::Matrix; I = -1, SOC = 1, tf = 1e6, kw...) -> (summary::Vector{Summary}, t, V, n_points)

`θ` holds one column per system in `theta_keys(h)` order (update_θ!, generate_functions.jl:364-372).  The keyword
names are those of `simulate` (model_evaluation.jl:17-44); hard failures of a single system re-throw the reference's
errors (model_evaluation.jl:456, checks.jl:233-239), in a batch they are per-system negative flags.
"""
function simulate_batch(h, θ::Matrix{Float64}; SOC = 1.0, tf = 1e6, nsave = 512, reltol = 1e-3, abstol = 1e-6,
                        bounds::Bounds = bounds_defaults(h), inputs...)
    length(inputs) == 1 || error("Cannot select more than one input from: (I, V, P, dT, η_p)")
    (name, value), = pairs(inputs)
    B, N = size(θ, 2), nstates(h)
    run = Run(METHOD[name], INPUT_VALUE, Float64(value), Float64(tf), 1, 0)
    o = opts_defaults(h)
    opts = Opts(abstol, reltol, abstol, reltol, o.maxiters, o.check_bounds, o.interp_final, o.skip_alg_deriv)
    soc0 = fill(Float64(SOC), B)
    Y = zeros(N, B); YP = zeros(N, B); soc = zeros(B); t_end = zeros(B)
    summ = Vector{Summary}(undef, B)
    t = fill(NaN, nsave, B); V = fill(NaN, nsave, B); n = zeros(Cint, B)
    simulate!(h, B, θ, run, C_NULL, opts, bounds, soc0, Y, YP, soc, t_end, summ, nsave, t, V, C_NULL, C_NULL, C_NULL, C_NULL, n)
    if B == 1 && summ[1].flag < 0
        error(summ[1].flag == -1 ? "Could not initialize DAE in 100 iterations." : "Model failed to converge")
    end
    return summ, t, V, n
end

end # module
