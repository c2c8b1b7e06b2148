"""GPU parity tests of the kernels, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances: complex128 amplitudes <= 1e-12 max-abs, energies <= 1e-10 (north_star: 1e-10 Ha); integers bit-exact."""

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-12
E_TOL = 1e-10


@pytest.fixture(scope="module")
def DeviceState(native_lib):
    from qibochem_b200.engine import DeviceState as DS

    return DS


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return psi / np.linalg.norm(psi)


def random_string_masks(n, rng, kind="any"):
    """Random Pauli string as (x, z) masks."""
    if kind == "diag":
        x = 0
    elif kind == "low":
        x = int(rng.integers(1, min(32, 2**n)))
    elif kind == "high":
        x = int(rng.integers(1, 2**n)) | (1 << (n - 1))
    else:
        x = int(rng.integers(0, 2**n))
    z = int(rng.integers(0, 2**n))
    return x, z


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 9, 12])
def test_basis_state_and_copy(DeviceState, n):
    with DeviceState(n) as st:
        idx = (1 << n) - 1 if n < 3 else 5
        st.set_basis_state(idx)
        v = st.to_numpy()
        ref = np.zeros(2**n, dtype=complex)
        ref[idx] = 1
        assert np.array_equal(v, ref)
        assert st.norm2() == 1.0
        psi = random_state(n, n)
        st.from_numpy(psi)
        assert np.array_equal(st.to_numpy(), psi)
        assert abs(st.norm2() - 1.0) < 1e-13


@pytest.mark.parametrize("n", [1, 2, 4, 5, 6, 8, 11, 14])
@pytest.mark.parametrize("kind", ["any", "diag", "low", "high"])
def test_single_rotations_match_oracle(DeviceState, n, kind):
    rng = np.random.default_rng(100 * n + len(kind))
    psi = random_state(n, 7 * n + 1)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        ref = psi.copy()
        xs, zs, phis = [], [], []
        for _ in range(6):
            x, z = random_string_masks(n, rng, kind)
            phi = float(rng.uniform(-np.pi, np.pi))
            xs.append(x), zs.append(z), phis.append(phi)
            ref = O.apply_pauli_rotation(ref, x, z, phi)
        st.apply_pauli_rotations(xs, zs, phis)
        assert np.abs(st.to_numpy() - ref).max() < AMP_TOL


@pytest.mark.parametrize("n", [4, 6, 10, 13])
def test_fused_runs_match_sequential_oracle(DeviceState, n):
    """Runs of commuting same-x strings (like the 8 strings of a JW double) fuse into one pass."""
    rng = np.random.default_rng(n)
    psi = random_state(n, 3 * n)
    xs, zs, phis = [], [], []
    for _ in range(5):
        x = int(rng.integers(1, 2**n))
        z0 = int(rng.integers(0, 2**n))
        members = [z0]
        for _ in range(40):
            dz = int(rng.integers(0, 2**n)) & (x | int(rng.integers(0, 2**n)) & 0x7)
            if bin(x & dz).count("1") % 2 == 0 and (z0 ^ dz) not in members:
                members.append(z0 ^ dz)
            if len(members) == 8:
                break
        for z in members:
            xs.append(x), zs.append(z), phis.append(float(rng.uniform(-1, 1)))
    # and a diagonal run
    for _ in range(6):
        xs.append(0), zs.append(int(rng.integers(1, 2**n)) & 0x3F), phis.append(float(rng.uniform(-1, 1)))
    ref = psi.copy()
    for x, z, p in zip(xs, zs, phis):
        ref = O.apply_pauli_rotation(ref, x, z, p)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        passes = st.apply_pauli_rotations(xs, zs, phis)
        assert passes < len(xs)
        assert np.abs(st.to_numpy() - ref).max() < AMP_TOL


def test_ucc_double_is_one_pass_and_matches_gate_path(DeviceState):
    """HF + JW double excitation: mask path on the GPU == the reference's gate-by-gate program in the oracle."""
    n, ne, theta = 6, 2, 0.3173
    gates = O.hf_gates(n, ne) + O.ucc_gates([0, 1, 4, 5], theta) + O.ucc_gates([1, 3], -0.77)
    ref = O.run_gates(n, gates)
    xs, zs, phis = [], [], []
    for exc, th in (([0, 1, 4, 5], theta), ([1, 3], -0.77)):
        for string, coeff in O.ucc_pauli_terms(exc):
            x, z = O.pauli_masks(n, string)
            xs.append(x), zs.append(z), phis.append((2j * coeff * th).real)
    with DeviceState(n) as st:
        st.set_option("k1_mode", 1)
        st.set_basis_state(((1 << ne) - 1) << (n - ne))
        passes = st.apply_pauli_rotations(xs, zs, phis)
        assert passes == 2  # 8 strings of the double -> 1 pass, 2 strings of the single -> 1 pass
        assert np.abs(st.to_numpy() - ref).max() < AMP_TOL
        # default mode on a small register (n <= 13): the whole program is ONE launch with the state in shared memory
        st.set_option("k1_mode", 0)
        st.set_basis_state(((1 << ne) - 1) << (n - ne))
        assert st.apply_pauli_rotations(xs, zs, phis) == 1
        assert np.abs(st.to_numpy() - ref).max() < AMP_TOL


@pytest.mark.parametrize("n", [1, 2, 4, 5, 6, 9, 12, 15])
def test_expectation_matches_oracle(DeviceState, n):
    rng = np.random.default_rng(n)
    psi = random_state(n, 11 * n)
    T = 60
    xs, zs = [], []
    for k in range(T):
        kind = ["any", "diag", "low", "high"][k % 4]
        x, z = random_string_masks(n, rng, kind)
        if k % 5 == 0 and xs:
            x = xs[-1]  # repeated x-masks form groups
        xs.append(x), zs.append(z)
    cs = rng.normal(size=T)
    ref_terms = np.array([np.vdot(psi, O.apply_pauli(psi, x, z)).real for x, z in zip(xs, zs)])
    ref = float(np.dot(cs, ref_terms))
    with DeviceState(n) as st:
        st.from_numpy(psi)
        e = st.expectation(xs, zs, cs)
        assert abs(e - ref) < E_TOL
        assert st.last_sweeps <= len(set(xs))
        e2, per = st.expectation(xs, zs, cs, per_term=True)
        assert np.abs(per - ref_terms).max() < E_TOL
        assert abs(e2 - ref) < E_TOL


def test_expectation_large_group_is_split(DeviceState):
    n = 11
    rng = np.random.default_rng(5)
    psi = random_state(n, 2)
    zs = rng.choice(2**n, size=1500, replace=False)
    xs = np.zeros(1500, dtype=np.uint64)
    cs = rng.normal(size=1500)
    ref = O.expectation_masks(psi, xs, zs, cs)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        assert abs(st.expectation(xs, zs, cs) - ref) < E_TOL
        assert st.last_sweeps == 2


def test_h2_energies_from_doc_fixture(DeviceState):
    """Known answers of the reference docs (hamiltonian.rst:59, ansatz.rst:73,80) through K0/K1/K2."""
    from tests.fixtures import H2_JW_TERMS, masks_from_terms

    n = 4
    const, xs, zs, cs = masks_from_terms(n, H2_JW_TERMS)
    with DeviceState(n) as st:
        st.set_basis_state(0b1100)
        assert abs(const + st.expectation(xs, zs, cs) - (-1.11628373627429)) < 1e-13
        st.set_basis_state(0)
        assert abs(const + st.expectation(xs, zs, cs) - 0.7074183344740923) < 1e-13
        # UCCD scan reaches the FCI energy of the fixture
        rx, rz, rc = [], [], []
        for string, coeff in O.ucc_pauli_terms([0, 1, 2, 3]):
            x, z = O.pauli_masks(n, string)
            rx.append(x), rz.append(z), rc.append(coeff)
        best = 0.0
        for th in np.linspace(0.11, 0.12, 41):
            st.set_basis_state(0b1100)
            st.apply_pauli_rotations(rx, rz, [(2j * c * th).real for c in rc])
            best = min(best, const + st.expectation(xs, zs, cs))
        assert abs(best - (-1.1371622544371)) < 2e-7


@pytest.mark.parametrize("n_shots", [1, 31, 1000, 100_003])
def test_parity_sums_bit_exact(DeviceState, n_shots):
    rng = np.random.default_rng(n_shots)
    samples = rng.integers(0, 2**40, size=n_shots, dtype=np.uint64)
    masks = rng.integers(0, 2**40, size=300, dtype=np.uint64)
    masks[0] = 0
    ref = O.parity_sums(samples, masks)
    with DeviceState(3) as st:
        assert np.array_equal(st.parity_sums(samples, masks), ref)
        vals, counts = O.frequencies_from_samples(samples % np.uint64(64))
        ref_w = O.parity_sums(samples % np.uint64(64), masks)
        assert np.array_equal(st.parity_sums(vals, masks, counts=counts), ref_w)


@pytest.mark.parametrize("bits", [3, 14, 30])
def test_frequencies_bit_exact(DeviceState, bits):
    rng = np.random.default_rng(bits)
    samples = rng.integers(0, 2**bits, size=50_000, dtype=np.uint64) << np.uint64(2)
    keep = ((1 << bits) - 1) << 2
    vals, counts = O.frequencies_from_samples(samples >> np.uint64(2))
    with DeviceState(3) as st:
        k, c = st.frequencies(samples, keep)
        assert np.array_equal(k, vals) and np.array_equal(c, counts)


@pytest.mark.parametrize("n", [3, 7, 13, 16, 18])  # <= 16: full CDF + a thread per shot; 18: chunk sums + a warp per shot
def test_sampling_inverse_cdf_matches_oracle(DeviceState, n):
    """Same Philox uniforms + same probabilities => same indices (except shots that land within rounding of a
    bin edge), in Z, X and Y bases."""
    psi = random_state(n, 5 * n)
    rng = np.random.default_rng(n)
    xb = int(rng.integers(0, 2**n))
    yb = int(rng.integers(0, 2**n)) & ~xb
    basis = {n - 1 - b: ("X" if (xb >> b) & 1 else "Y") for b in range(n) if ((xb | yb) >> b) & 1}
    rot = O.run_gates(n, O.measurement_basis_gates(basis), psi)
    p = np.abs(rot) ** 2
    cdf = np.cumsum(p)
    n_shots, seed = 20_000, 1234 + n
    u = O.philox_uniform53(n_shots, seed)
    target = u * cdf[-1]
    ref_idx = np.searchsorted(cdf, target, side="right")
    with DeviceState(n) as st:
        st.from_numpy(psi)
        got, total = st.sample(xb, yb, n_shots, seed, return_total=True)
        assert abs(total - 1.0) < 1e-12
        assert np.array_equal(st.to_numpy(), psi)  # psi untouched
        edge = np.minimum(np.abs(cdf[np.minimum(ref_idx, 2**n - 1)] - target), np.abs(cdf[np.maximum(ref_idx - 1, 0)] - target))
        mismatch = got != ref_idx.astype(np.uint64)
        assert np.all(edge[mismatch] < 1e-12)
        assert mismatch.mean() < 1e-3
        # device-resident samples feed K3 without a host round trip
        dev = st.sample(xb, yb, n_shots, seed, on_device=True)
        masks = rng.integers(0, 2**n, size=17, dtype=np.uint64)
        assert np.array_equal(st.parity_sums(dev, masks), O.parity_sums(got, masks))


@pytest.mark.parametrize("bit", [0, 3, 7, 9])
def test_global_local_bit_swap_pair(DeviceState, bit):
    """K4 on one GPU: two shards of an (n+1)-qubit state, swap the global bit with local bit `bit`."""
    n = 10
    full = random_state(n + 1, 99)
    with DeviceState(n) as lo, DeviceState(n) as hi:
        lo.from_numpy(full[: 2**n])
        hi.from_numpy(full[2**n :])
        lo.swap_local_pair(hi, bit)
        idx = np.arange(2 ** (n + 1))
        g = (idx >> n) & 1
        l = (idx >> bit) & 1
        src = (idx & ~((1 << n) | (1 << bit))) | (l << n) | (g << bit)
        ref = full[src]
        got = np.concatenate([lo.to_numpy(), hi.to_numpy()])
        assert np.array_equal(got, ref)
        # the two-rank form (each side moves half of the pairs) composes to the same permutation
        lo.from_numpy(full[: 2**n])
        hi.from_numpy(full[2**n :])
        lo.swap_with_peer(hi.device_ptr, bit, 0, 0)
        hi.swap_with_peer(lo.device_ptr, bit, 1, 1)
        lo.synchronize(), hi.synchronize()
        got = np.concatenate([lo.to_numpy(), hi.to_numpy()])
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("n,T", [(12, 400), (13, 1200), (16, 3000)])
def test_tiled_expectation_matches_sweeps_and_oracle(DeviceState, n, T):
    """K2 tiled (shared-memory tiles + register sub-cubes) == K2 sweeps == oracle, on a molecular-structured term
    list plus terms that exercise every branch: odd-nY (Im q), weight-1/3/5 x-masks, x-masks too wide for a tile."""
    from qibochem_b200 import synthetic

    rng = np.random.default_rng(n)
    psi = random_state(n, 31 * n)
    hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(n, T)
    extra_x, extra_z = [], []
    for w in (1, 1, 2, 3, 3, 5, 5, 6, 7, 4, 4, 2):
        bits = rng.choice(n, size=w, replace=False)
        x = int(sum(1 << int(b) for b in bits))
        for _ in range(3):
            extra_x.append(x), extra_z.append(int(rng.integers(0, 2**n)))
    hx = np.concatenate([hx, np.array(extra_x, dtype=np.uint64)])
    hz = np.concatenate([hz, np.array(extra_z, dtype=np.uint64)])
    hc = np.concatenate([hc, rng.normal(size=len(extra_x))])
    ref = O.expectation_masks(psi, hx.tolist(), hz.tolist(), hc.tolist())
    with DeviceState(n) as st:
        st.from_numpy(psi)
        st.set_option("k2_mode", 1)
        e_sweeps = st.expectation(hx, hz, hc)
        sweeps = st.last_sweeps
        st.set_option("k2_mode", 2)
        e_tiled = st.expectation(hx, hz, hc)
        passes = st.last_sweeps
        e_tiled_again = st.expectation(hx, hz, hc)  # cached plan
        assert abs(e_sweeps - ref) < E_TOL
        assert abs(e_tiled - ref) < E_TOL
        assert e_tiled == e_tiled_again  # deterministic reduction order
        assert passes < sweeps
        # a different coefficient vector must not reuse the cached tables
        e2 = st.expectation(hx, hz, 2.0 * hc)
        assert abs(e2 - 2.0 * ref) < 2 * E_TOL


def test_tiled_expectation_on_ucc_state(DeviceState):
    """Physically shaped input: HF + UCCSD prefix state (sparse, particle-number conserving) at 18 qubits."""
    from qibochem_b200 import synthetic

    n, ne = 18, 8
    rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, ne, max_excitations=40)
    hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, 8000)
    with DeviceState(n) as st:
        st.set_basis_state(synthetic.hf_index(n, ne))
        st.apply_pauli_rotations(rx, rz, ra)
        st.set_option("k2_mode", 1)
        e1 = st.expectation(hx, hz, hc)
        st.set_option("k2_mode", 2)
        e2 = st.expectation(hx, hz, hc)
        st.set_option("k2_mode", 0)
        e0 = st.expectation(hx, hz, hc)
        psi = st.to_numpy()
    from oracle import oracle_c as OC

    ref = OC.expectation_masks(psi, n, hx, hz, hc)
    assert abs(e1 - ref) < E_TOL and abs(e2 - ref) < E_TOL and abs(e0 - ref) < E_TOL


@pytest.mark.parametrize("n", [12, 13, 15, 16])
def test_tiled_rotation_batches_match_oracle(DeviceState, n):
    """K1T: consecutive runs whose x-masks fit a 12-bit tile are applied in one in-place pass.  Random fused runs with
    z / support bits inside AND outside the tile, diagonal runs, low-x runs; tiles == one-pass-per-run == oracle."""
    rng = np.random.default_rng(50 + n)
    psi = random_state(n, 5 * n + 2)
    xs, zs, phis = [], [], []
    pool = rng.choice(n, size=7, replace=False)  # x-masks drawn from 7 bits: batches form
    for run in range(14):
        if run % 5 == 4:
            x = 0
        else:
            w = int(rng.integers(1, 5))
            x = int(sum(1 << int(b) for b in rng.choice(pool, size=w, replace=False)))
        z0 = int(rng.integers(0, 2**n))
        members = [z0]
        for _ in range(30):
            dz = int(rng.integers(0, 2**n)) & (x | (int(rng.integers(0, 2**n)) & (0x3 | 1 << (n - 1) | 1 << (n - 2))))
            if bin(x & dz).count("1") % 2 == 0 and (z0 ^ dz) not in members:
                members.append(z0 ^ dz)
            if len(members) == 6:
                break
        for z in members:
            xs.append(x), zs.append(z), phis.append(float(rng.uniform(-1, 1)))
    # a string too wide for a tile in the middle of the stream
    xs.insert(20, (1 << n) - 1), zs.insert(20, 5), phis.insert(20, 0.3)
    ref = psi.copy()
    for x, z, p in zip(xs, zs, phis):
        ref = O.apply_pauli_rotation(ref, x, z, p)
    with DeviceState(n) as st:
        results = {}
        for mode in (1, 0, 2, 3):
            st.set_option("k1_mode", mode)
            st.from_numpy(psi)
            passes = st.apply_pauli_rotations(xs, zs, phis)
            results[mode] = (passes, st.to_numpy())
            assert np.abs(results[mode][1] - ref).max() < AMP_TOL, f"k1_mode={mode}"
        assert results[0][0] < results[1][0] and results[2][0] < results[1][0]  # batching saved passes


def test_tiled_rotations_on_uccsd_stream(DeviceState):
    """UCCSD prefix at 16 qubits (reference excitation order): batched passes == one pass per excitation == C oracle."""
    from oracle import oracle_c as OC
    from qibochem_b200 import synthetic

    n, ne = 16, 6
    rx, rz, ra, n_exc = synthetic.uccsd_rotation_stream(n, ne, max_excitations=60)
    ref = np.zeros(2**n, dtype=np.complex128)
    ref[synthetic.hf_index(n, ne)] = 1.0
    for x, z, a in zip(rx.tolist(), rz.tolist(), ra.tolist()):
        OC.pauli_rotation(ref, n, x, z, a)
    with DeviceState(n) as st:
        out = {}
        for mode in (1, 0, 2):
            st.set_option("k1_mode", mode)
            st.set_basis_state(synthetic.hf_index(n, ne))
            out[mode] = (st.apply_pauli_rotations(rx, rz, ra), st.to_numpy())
            assert np.abs(out[mode][1] - ref).max() < AMP_TOL
        assert out[1][0] == n_exc and out[0][0] * 2 < n_exc and out[2][0] * 2 < n_exc


def test_full_size_30q_properties(DeviceState):
    """BASELINE configs[3] size (30 qubits, 17 GB state): size-independent properties instead of an oracle run.
    (a) unitarity: norm stays 1; (b) known answer: a Z-only Hamiltonian on a basis state; (c) tile-batched rotations ==
    one pass per excitation; (d) tiled expectation == per-x-mask sweeps on a term subset; (e) linearity in the
    coefficients; (f) U^-1 U = 1: the reversed stream with negated angles returns the HF determinant."""
    import torch

    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs > 40 GB of HBM")
    from qibochem_b200 import synthetic

    n, ne = 30, 14
    rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, ne, max_excitations=24)
    hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(n, 100_000)
    hf = synthetic.hf_index(n, ne)
    rng = np.random.default_rng(30)
    sub = np.sort(rng.choice(hx.size, size=48, replace=False))
    big = np.sort(rng.choice(hx.size, size=12_000, replace=False))
    with DeviceState(n) as st:
        # (b) diagonal known answer on |HF>
        st.set_basis_state(hf)
        diag = hx == 0
        known = float(np.sum(hc[diag] * np.where(np.array([bin(hf & int(z)).count("1") & 1 for z in hz[diag]]) == 1, -1.0, 1.0)))
        assert abs(st.expectation(hx[diag], hz[diag], hc[diag]) - known) < E_TOL
        # (c) batched vs per-run rotations, (a) norm
        energies = {}
        for mode in (1, 0):
            st.set_option("k1_mode", mode)
            st.set_basis_state(hf)
            passes = st.apply_pauli_rotations(rx, rz, ra)
            assert abs(st.norm2() - 1.0) < 1e-12
            st.set_option("k2_mode", 1)
            energies[mode] = (passes, st.expectation(hx[sub], hz[sub], hc[sub]))
        assert energies[0][0] < energies[1][0]
        assert abs(energies[0][1] - energies[1][1]) < E_TOL
        # (d) tiles vs sweeps on the same subset (state: the batched one)
        st.set_option("k2_mode", 2)
        assert abs(st.expectation(hx[sub], hz[sub], hc[sub]) - energies[0][1]) < E_TOL
        # (e) linearity on a 12k-term subset through the tiled path
        st.set_option("k2_mode", 0)
        c1 = hc[big]
        c2 = rng.normal(scale=0.05, size=big.size)
        e1 = st.expectation(hx[big], hz[big], c1)
        e2 = st.expectation(hx[big], hz[big], c2)
        e12 = st.expectation(hx[big], hz[big], c1 + 2.0 * c2)
        assert abs(e12 - (e1 + 2.0 * e2)) < 1e-9
        # (f) inverse stream
        st.apply_pauli_rotations(rx[::-1], rz[::-1], -ra[::-1])
        amp = st.to_numpy(offset=hf, count=1)[0]
        assert abs(amp - 1.0) < 1e-12 and abs(st.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("n", [2, 3, 6, 10, 13])
def test_adjoint_gradient_matches_parameter_shift_oracle(DeviceState, n):
    """qc_energy_gradient (lambda = H psi + backward walk over the fused runs) == the parameter-shift rule evaluated
    with the numpy oracle, for a program with fused runs, diagonal and low-x strings and odd-nY Hamiltonian terms."""
    rng = np.random.default_rng(900 + n)
    xs, zs, phis = [], [], []
    for run in range(5):
        kind = ["any", "diag", "low", "high", "any"][run]
        x, z0 = random_string_masks(n, rng, kind)
        members = [z0]
        for _ in range(20):
            dz = int(rng.integers(0, 2**n))
            if bin(x & dz).count("1") % 2 == 0 and (z0 ^ dz) not in members:
                members.append(z0 ^ dz)
            if len(members) == 3:
                break
        for z in members:
            xs.append(x), zs.append(z), phis.append(float(rng.uniform(-1, 1)))
    T = 40
    hx = [int(rng.integers(0, 2**n)) if k % 3 else 0 for k in range(T)]
    hx[5] = hx[6] = hx[7]  # a group with several z-masks
    hz = [int(rng.integers(0, 2**n)) for _ in range(T)]
    hc = rng.normal(size=T)
    psi0 = random_state(n, 17 * n)

    def energy(angles):
        v = psi0.copy()
        for x, z, p in zip(xs, zs, angles):
            v = O.apply_pauli_rotation(v, x, z, p)
        return O.expectation_masks(v, hx, hz, hc.tolist())

    ref_e = energy(phis)
    ref_g = np.zeros(len(phis))
    for k in range(len(phis)):
        up, dn = list(phis), list(phis)
        up[k] += np.pi / 2
        dn[k] -= np.pi / 2
        ref_g[k] = 0.5 * (energy(up) - energy(dn))
    with DeviceState(n) as st:
        st.from_numpy(psi0)
        st.apply_pauli_rotations(xs, zs, phis)
        e, g = st.energy_gradient(xs, zs, phis, hx, hz, hc)
        assert abs(e - ref_e) < E_TOL
        assert np.abs(g - ref_g).max() < 1e-10
        assert np.abs(st.to_numpy() - psi0).max() < AMP_TOL  # the program has been un-applied
        # and the tiled / sweep expectation agrees with <psi|H psi> of the adjoint pass
        st.apply_pauli_rotations(xs, zs, phis)
        assert abs(st.expectation(hx, hz, hc) - e) < E_TOL


@pytest.mark.parametrize("n,seed", [(12, 0), (14, 1), (15, 2)])
def test_tiled_expectation_random_dense_hamiltonian(DeviceState, n, seed):
    """Unstructured term lists: random x-masks of every weight 0..7 with random z-masks, large same-x groups with
    many distinct z (pattern sums + LIN entries, split blobs), duplicates; tiles == oracle."""
    from oracle import oracle_c as OC

    rng = np.random.default_rng(seed)
    xs, zs = [], []
    for _ in range(2500):
        w = int(rng.integers(0, 8))
        x = int(sum(1 << int(b) for b in rng.choice(n, size=w, replace=False)))
        xs.append(x), zs.append(int(rng.integers(0, 2**n)))
    for x in (0, 0b11, 0b10110, (1 << (n - 1)) | 1):  # big groups: >> 300 distinct z per x-mask
        for z in rng.choice(2**n, size=900, replace=False):
            xs.append(x), zs.append(int(z))
    xs += xs[:50]
    zs += zs[:50]  # duplicated terms simply add up
    cs = rng.normal(size=len(xs))
    psi = random_state(n, 99 + seed)
    ref = OC.expectation_masks(psi, n, xs, zs, cs)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        st.set_option("k2_mode", 2)
        e = st.expectation(xs, zs, cs)
        assert abs(e - ref) < 1e-9 * max(1.0, np.abs(cs).sum() / 50)
        st.set_option("k2_mode", 1)
        assert abs(st.expectation(xs, zs, cs) - e) < 1e-9 * max(1.0, np.abs(cs).sum() / 50)


@pytest.mark.parametrize("n,seed", [(13, 0), (16, 1), (18, 2)])
def test_tiled_expectation_hopping_groups(DeviceState, n, seed):
    """Weight-2 x-masks (the HOP record path of the tiled expectation and its fall-backs): for random pairs {p, q}
    the strings X..X / Y..Y with a shared random Z background, decorated with one or two extra Z (hop-able: two real
    patterns, many sign-only subgroups), plus groups with X..Y / Y..X strings (odd nY: imaginary patterns), with three
    Y-patterns and with a handful of terms only (the generic record path); tiles == sweeps == oracle."""
    from oracle import oracle_c as OC

    rng = np.random.default_rng(100 + seed)
    bit = lambda q: 1 << q  # noqa: E731
    xs, zs = [], []
    pairs = [tuple(sorted(rng.choice(n, size=2, replace=False).tolist())) for _ in range(60)]
    for k, (p, q) in enumerate(dict.fromkeys(pairs)):
        x = bit(p) | bit(q)
        others = [r for r in range(n) if r not in (p, q)]
        back = int(sum(bit(r) for r in others if rng.random() < 0.4))  # shared Z background (a "chain")
        ys = [0, x] if k % 4 else [0, x, bit(p), bit(q)]  # every fourth group also has XY / YX strings
        if k % 7 == 3:
            ys = [0]  # a single pattern
        decos = [0] + [bit(r) for r in others] + [bit(a) | bit(b) for a, b in zip(others[::2], others[1::2])]
        if k % 5 == 4:
            decos = decos[:3]  # too few subgroups for pattern sums
        for y in ys:
            for d in decos:
                xs.append(x), zs.append(back ^ y ^ d)
    cs = rng.normal(size=len(xs))
    psi = random_state(n, 7 + seed)
    ref = OC.expectation_masks(psi, n, xs, zs, cs)
    tol = 1e-10 * max(1.0, np.abs(cs).sum() / 50)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        st.set_option("k2_mode", 2)
        e = st.expectation(xs, zs, cs)
        assert abs(e - ref) < tol
        assert e == st.expectation(xs, zs, cs)  # deterministic
        st.set_option("k2_mode", 1)
        assert abs(st.expectation(xs, zs, cs) - ref) < tol


def test_empty_and_degenerate_inputs(DeviceState):
    """Empty programs / term lists, tile modes requested on registers too small for a tile, repeated identical calls."""
    n = 6
    psi = random_state(n, 4)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        assert st.apply_pauli_rotations([], [], []) == 0
        assert st.expectation([], [], []) == 0.0
        e, g = st.energy_gradient([], [], [], [3], [5], [0.5])
        assert g.size == 0 and abs(e - O.expectation_masks(psi, [3], [5], [0.5])) < E_TOL
        e, g = st.energy_gradient([1], [2], [0.3], [], [], [])
        assert e == 0.0 and g.tolist() == [0.0]
        st.set_option("k2_mode", 2)  # tiles need >= 12 local qubits: falls back to sweeps
        st.set_option("k1_mode", 3)
        xs, zs, cs = [0, 7, 7, 40], [5, 1, 6, 9], [0.3, -0.2, 0.9, 1.5]
        assert abs(st.expectation(xs, zs, cs) - O.expectation_masks(psi, xs, zs, cs)) < E_TOL
        st.apply_pauli_rotations([7, 7], [1, 6], [0.4, -0.1])
        ref = O.apply_pauli_rotation(O.apply_pauli_rotation(psi, 7, 1, 0.4), 7, 6, -0.1)
        assert np.abs(st.to_numpy() - ref).max() < AMP_TOL
        with pytest.raises(RuntimeError):
            st.set_option("no_such_option", 1)
        with pytest.raises(RuntimeError):
            st.expectation([1 << n], [0], [1.0])  # mask outside the register


@pytest.mark.parametrize("n", [10, 11, 14])
def test_register_phase_rotations_edge_cases(DeviceState, n):
    """K1O phases: repeated x-masks, x-masks in the span of earlier ones (phase break), more than five distinct masks,
    low / high / diagonal runs mixed, z-masks with bits everywhere; register phases == one pass per run == oracle."""
    rng = np.random.default_rng(77 + n)
    psi = random_state(n, 3 * n + 1)
    b = [1 << int(k) for k in rng.permutation(n)[:7]]
    xmasks = [b[0] | b[1], b[2] | b[3], b[0] | b[1], 0, b[0] | b[1] | b[2] | b[3],  # the last one is x0 ^ x1: dependent
              b[4], b[5], b[6], b[4] | b[5] | b[6], b[0], 1, 3, 2, 0, (1 << (n - 1)) | 1, b[2] | b[3]]  # fmt: skip
    xs, zs, phis = [], [], []
    for x in xmasks:
        z0 = int(rng.integers(0, 2**n))
        members = [z0]
        for _ in range(12):
            dz = int(rng.integers(0, 2**n)) & (x | 0x5)
            if bin(x & dz).count("1") % 2 == 0 and (z0 ^ dz) not in members:
                members.append(z0 ^ dz)
            if len(members) == 3:
                break
        for z in members:
            xs.append(x), zs.append(z), phis.append(float(rng.uniform(-1.5, 1.5)))
    ref = psi.copy()
    for x, z, p in zip(xs, zs, phis):
        ref = O.apply_pauli_rotation(ref, x, z, p)
    with DeviceState(n) as st:
        for mode in (3, 0, 1):
            st.set_option("k1_mode", mode)
            st.from_numpy(psi)
            st.apply_pauli_rotations(xs, zs, phis)
            assert np.abs(st.to_numpy() - ref).max() < AMP_TOL, f"k1_mode={mode}"


def test_tiled_plan_cache_compares_term_lists(DeviceState):
    """The tiled plan bakes the coefficients in and is cached per handle: H, -H and H with two signs flipped used to
    hash alike (word-wise xor-multiply keeps bit 63 in bit 63) and silently reuse H's plan."""
    from qibochem_b200 import synthetic

    n = 16
    psi = random_state(n, 77)
    hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(n, 2000)
    if hc.size % 2:
        hx, hz, hc = hx[:-1], hz[:-1], hc[:-1]
    flipped = hc.copy()
    flipped[[3, 700]] *= -1.0
    scaled = hc * 4.0  # exponent-only change
    with DeviceState(n) as st:
        st.from_numpy(psi)
        st.set_option("k2_mode", 1)
        ref = {name: st.expectation(hx, hz, c) for name, c in (("h", hc), ("neg", -hc), ("flip", flipped), ("x4", scaled))}
        st.set_option("k2_mode", 0)  # n >= 16, T >= 512: tiles
        got = {name: st.expectation(hx, hz, c) for name, c in (("h", hc), ("neg", -hc), ("flip", flipped), ("x4", scaled), ("h2", hc))}
    assert abs(ref["neg"] + ref["h"]) < E_TOL and abs(ref["flip"] - ref["h"]) > 1e-6
    for name in ("h", "neg", "flip", "x4"):
        assert abs(got[name] - ref[name]) < E_TOL, name
    assert got["h2"] == got["h"]


def test_tiled_expectation_with_many_leftover_groups(DeviceState):
    """More leftover x-groups (too wide for a tile) than the result buffer's initial capacity: growing it while the
    tiled sum already sits in slot 0 must keep that sum."""
    from qibochem_b200 import synthetic

    n = 17
    rng = np.random.default_rng(5)
    psi = random_state(n, 78)
    hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(n, 1500)
    from itertools import combinations

    all_wide = [sum(1 << b for b in c) for c in combinations(range(n), 13)]  # C(17, 13) = 2380 masks wider than a tile
    wx = np.array(sorted(rng.choice(all_wide, size=1200, replace=False).tolist()), dtype=np.uint64)
    wz = rng.integers(0, 2**n, size=wx.size, dtype=np.uint64)
    wc = rng.normal(scale=0.05, size=wx.size)
    ax, az, ac = np.concatenate([hx, wx]), np.concatenate([hz, wz]), np.concatenate([hc, wc])
    with DeviceState(n) as st:  # fresh handle: result buffer at its initial capacity
        st.from_numpy(psi)
        st.set_option("k2_mode", 2)
        e_tiled = st.expectation(ax, az, ac)
        st.set_option("k2_mode", 1)
        e_sweeps = st.expectation(ax, az, ac)
    from oracle import oracle_c as OC

    ref = OC.expectation_masks(psi, n, ax, az, ac)
    assert abs(e_sweeps - ref) < E_TOL and abs(e_tiled - ref) < E_TOL


def test_bench_generators_26q_against_oracle(DeviceState):
    """The benchmark's own generators (UCCSD prefix + molecular-structured Hamiltonian; 76 000 terms is all a 26-qubit
    register offers: every 4-set is used) at 26 qubits: per-term values of a random 1000-term subset against the C
    oracle on the SAME state (copied back), the tiled energy of that subset against the oracle sum, and the full
    tiled energy against the per-x-mask sweeps."""
    import torch

    if torch.cuda.get_device_properties(0).total_memory < 8e9:
        pytest.skip("needs > 8 GB of HBM")
    from oracle import oracle_c as OC
    from qibochem_b200 import synthetic

    n, ne = 26, 12
    OC.set_num_threads(0)
    rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, ne, max_excitations=64)
    hx, hz, hc, const = synthetic.molecular_like_hamiltonian(n, 76_000)
    n_sub = 1000 if OC.num_threads() >= 12 else 300
    sub = np.sort(np.random.default_rng(26).choice(hx.size, size=n_sub, replace=False))
    with DeviceState(n) as st:
        st.set_basis_state(synthetic.hf_index(n, ne))
        st.apply_pauli_rotations(rx, rz, ra)
        e_sub_pt, per_term = st.expectation(hx[sub], hz[sub], hc[sub], per_term=True)
        st.set_option("k2_mode", 2)
        e_sub_tiled = st.expectation(hx[sub], hz[sub], hc[sub])
        e_full_tiled = st.expectation(hx, hz, hc)
        passes = st.last_sweeps
        st.set_option("k2_mode", 1)
        e_full_sweeps = st.expectation(hx, hz, hc)
        sweeps = st.last_sweeps
        psi = st.to_numpy()
    assert abs(np.vdot(psi, psi).real - 1.0) < 1e-12
    ref_terms = OC.expectation_masks_per_term(psi, n, hx[sub], hz[sub])
    assert np.abs(per_term - ref_terms).max() < 1e-12
    ref_sub = float(np.dot(hc[sub], ref_terms))
    assert abs(e_sub_pt - ref_sub) < E_TOL and abs(e_sub_tiled - ref_sub) < E_TOL
    assert passes < sweeps / 20
    assert abs(e_full_tiled - e_full_sweeps) < E_TOL


def _hpsi_inputs(n, seed):
    """Hamiltonian for the tiled H|psi>: molecular-structured terms (quads, hopping groups with ~2n z-masks each, the
    diagonal group) + random strings with odd nY (Im-type weights), several z-masks per x-mask, and x-masks too wide
    for a 12-bit tile (sweep fall-back)."""
    from qibochem_b200 import synthetic

    rng = np.random.default_rng(seed)
    hx, hz, hc, _ = synthetic.molecular_like_hamiltonian(n, 900)
    ex, ez = [], []
    for w in (1, 2, 3, 3, 5, 6, 4, 13 if n >= 13 else 4, 7, 2):
        bits = rng.choice(n, size=min(w, n), replace=False)
        x = int(sum(1 << int(b) for b in bits))
        for _ in range(3):
            ex.append(x), ez.append(int(rng.integers(0, 2**n)))
    hx = np.concatenate([hx, np.array(ex, dtype=np.uint64)])
    hz = np.concatenate([hz, np.array(ez, dtype=np.uint64)])
    hc = np.concatenate([hc, rng.normal(size=len(ex))])
    return hx, hz, hc


@pytest.mark.parametrize("n", [12, 13, 15, 16])
def test_tiled_hpsi_matches_oracle(DeviceState, n):
    """lambda = H psi through tiles (K6T) == the oracle's sum of Pauli applications, amplitude by amplitude, in both
    the overwrite and the accumulate mode; <psi|lambda> == the energy of the sweep kernel."""
    hx, hz, hc = _hpsi_inputs(n, 40 + n)
    psi = random_state(n, 3 * n)
    ref = np.zeros_like(psi)
    for x, z, c in zip(hx.tolist(), hz.tolist(), hc.tolist()):
        ref += c * O.apply_pauli(psi, int(x), int(z))
    with DeviceState(n) as st:
        st.from_numpy(psi)
        st.set_option("k2_mode", 1)
        e_sweeps = st.expectation(hx, hz, hc)
        st.set_option("k2_mode", 2)
        st.hpsi(hx, hz, hc)
        ip = st.lambda_inner()
        assert abs(np.vdot(ref, psi) - ip) < 1e-11
        assert abs(ip.real - e_sweeps) < E_TOL
        half = hx.size // 2
        st.hpsi(hx[:half], hz[:half], hc[:half])
        st.hpsi(hx[half:], hz[half:], hc[half:], accumulate=True)
        assert abs(st.lambda_inner() - ip) < 1e-11
        # amplitude-level check of lambda through the gradient of a one-rotation program: dE/dphi = Im<lambda|P|psi>
        x0, z0 = int(hx[-1]), int(hz[-1])
        grad = st.gradient_backward([x0], [z0], [0.0])
        assert abs(grad[0] - np.vdot(ref, O.apply_pauli(psi, x0, z0)).imag) < 1e-11
        assert np.abs(st.to_numpy() - psi).max() < AMP_TOL


@pytest.mark.parametrize("n", [14, 16])
def test_adjoint_gradient_through_tiles(DeviceState, n):
    """The whole adjoint gradient with the tiled H|psi> on a UCCSD prefix + molecular-structured Hamiltonian ==
    the parameter-shift rule evaluated with the device's own (oracle-checked) energy path on a few parameters, and
    == the sweep-based adjoint gradient on all of them."""
    from qibochem_b200 import synthetic

    ne = 6
    rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, ne, max_excitations=6)
    hx, hz, hc = _hpsi_inputs(n, 50 + n)
    hf = synthetic.hf_index(n, ne)
    with DeviceState(n) as st:
        def energy(angles, mode):
            st.set_option("k2_mode", mode)
            st.set_basis_state(hf)
            st.apply_pauli_rotations(rx, rz, angles)
            return st.expectation(hx, hz, hc)

        grads = {}
        for mode in (1, 2):
            e0 = energy(ra, mode)
            e, grads[mode] = st.energy_gradient(rx, rz, ra, hx, hz, hc)
            assert abs(e - e0) < E_TOL
            assert abs(st.to_numpy(offset=hf, count=1)[0] - 1.0) < AMP_TOL  # back at |HF>
        assert np.abs(grads[1] - grads[2]).max() < 1e-11
        for k in (0, 7, len(ra) - 1):
            up, dn = ra.copy(), ra.copy()
            up[k] += np.pi / 2
            dn[k] -= np.pi / 2
            assert abs(grads[2][k] - 0.5 * (energy(up, 1) - energy(dn, 1))) < 1e-10


@pytest.mark.parametrize("n", [13, 16, 18])
def test_real_amplitude_fast_path(DeviceState, n):
    """Option k2_real: a HF determinant evolved by UCC rotations (every string has an odd number of Y) is exactly real,
    the tiled expectation then runs the real-amplitude kernels: same energy as the complex kernels and the oracle,
    odd-nY Hamiltonian terms included (they contribute exactly 0 on a real state).  A complex state must NOT take
    the path."""
    from oracle import oracle_c as OC
    from qibochem_b200 import synthetic

    ne = 6
    rx, rz, ra, _ = synthetic.uccsd_rotation_stream(n, ne, max_excitations=30)
    hx, hz, hc = _hpsi_inputs(n, 60 + n)  # molecular-structured + random strings (odd nY among them) + a wide x-mask
    with DeviceState(n) as st:
        st.set_basis_state(synthetic.hf_index(n, ne))
        st.apply_pauli_rotations(rx, rz, ra)
        psi = st.to_numpy()
        assert np.abs(psi.imag).max() == 0.0  # exactly real, not merely small
        st.set_option("k2_mode", 2)
        e_complex = st.expectation(hx, hz, hc)
        assert not st.took_real_path()
        st.set_option("k2_real", 1)
        e_real = st.expectation(hx, hz, hc)
        assert st.took_real_path()
        ref = OC.expectation_masks(psi, n, hx, hz, hc)
        assert abs(e_complex - ref) < E_TOL and abs(e_real - ref) < E_TOL
        # a state with imaginary parts: the check fails and the complex kernels run
        st.apply_pauli_rotations([0], [1], [0.3])  # exp(-i 0.15 Z_{n-1}): diagonal phase
        e_c = st.expectation(hx, hz, hc)
        assert not st.took_real_path()
        st.set_option("k2_real", 0)
        assert abs(st.expectation(hx, hz, hc) - e_c) < 1e-13


@pytest.mark.parametrize("n,n_rot", [(5, 2), (12, 4), (14, 6), (14, 7)])
def test_sampling_small_register_fused_basis_change(DeviceState, n, n_rot):
    """Small registers (<= 16 qubits) with <= 6 rotated qubits compute |<i|B|psi>|^2 in one launch straight from psi
    (no scratch copy, no gate passes); 7 rotated qubits take the gate path.  Same Philox uniforms + the oracle's
    probabilities => same indices, except shots within rounding of a bin edge."""
    psi = random_state(n, 7 * n + n_rot)
    rng = np.random.default_rng(100 * n + n_rot)
    bits = rng.choice(n, size=n_rot, replace=False)
    xb = int(sum(1 << int(b) for b in bits[: n_rot // 2]))
    yb = int(sum(1 << int(b) for b in bits[n_rot // 2 :]))
    basis = {n - 1 - b: ("X" if (xb >> b) & 1 else "Y") for b in range(n) if ((xb | yb) >> b) & 1}
    rot = O.run_gates(n, O.measurement_basis_gates(basis), psi.copy())
    cdf = np.cumsum(np.abs(rot) ** 2)
    n_shots, seed = 30_000, 99 + n
    target = O.philox_uniform53(n_shots, seed) * cdf[-1]
    ref_idx = np.minimum(np.searchsorted(cdf, target, side="right"), 2**n - 1)
    with DeviceState(n) as st:
        st.from_numpy(psi)
        got, total = st.sample(xb, yb, n_shots, seed, return_total=True)
        assert abs(total - 1.0) < 1e-12
        assert np.array_equal(st.to_numpy(), psi)
    edge = np.minimum(np.abs(cdf[ref_idx] - target), np.abs(cdf[np.maximum(ref_idx - 1, 0)] - target))
    mismatch = got != ref_idx.astype(np.uint64)
    assert np.all(edge[mismatch] < 1e-12)
    assert mismatch.mean() < 1e-3
