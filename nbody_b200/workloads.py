"""Synthetic initial conditions, reproducible at any N.

Every random number is a pure function of (seed, particle index, stream):
``u = float(splitmix64(seed + GOLDEN*(8*i + k)) >> 40) * 2**-24`` so the same
particle set can be regenerated anywhere (host, GPU box, oracle) without files.
Particles are returned in the boundary layout of the reference's ``Particle``
(include/nbody/simulation.h:16-35 instantiated with a 16-byte float4 vector):
float32 [N, 12] = position[4], velocity[4], mass, charge, 8 bytes padding.
The uniform cube follows the reference demo driver's recipe (src/main.cpp:23-58).
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
PARTICLE_FLOATS = 12  # 48-byte AoS record


def _splitmix64(x):
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def uniform01(seed, idx, k):
    """float32 in [0,1) for particle indices ``idx`` (uint64 array) and stream k (0..7)."""
    with np.errstate(over="ignore"):
        ctr = np.uint64(seed) + GOLDEN * (np.uint64(8) * idx + np.uint64(k))
        return (_splitmix64(ctr) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def _pack(pos, vel, mass, charge):
    n = pos.shape[0]
    P = np.zeros((n, PARTICLE_FLOATS), np.float32)
    P[:, 0:3] = pos
    P[:, 4:7] = vel
    P[:, 8] = mass
    P[:, 9] = charge
    return P


def _isotropic(u_theta, u_phi):
    theta = np.float32(2.0 * np.pi) * u_theta
    cphi = np.float32(2.0) * u_phi - np.float32(1.0)
    sphi = np.sqrt(np.maximum(np.float32(0.0), np.float32(1.0) - cphi * cphi))
    return np.stack([sphi * np.cos(theta), sphi * np.sin(theta), cphi], axis=1).astype(np.float32)


def uniform_cube(n, seed=42, start=0):
    """Uniform random cube in [0,1)^3 (src/main.cpp:23-58): |v| = 0.1 isotropic,
    mass in [1,10), charge = mass (gravity). ``start`` offsets the particle index
    so ranks can generate disjoint slices of one global set."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    pos = np.stack([uniform01(seed, idx, k) for k in range(3)], axis=1)
    vel = np.float32(0.1) * _isotropic(uniform01(seed, idx, 3), uniform01(seed, idx, 4))
    mass = np.float32(1.0) + np.float32(9.0) * uniform01(seed, idx, 5)
    return _pack(pos, vel, mass, mass)


def _plummer(n, seed, start, centre, scale, rmax, total_mass, bulk_v, n_total):
    idx = np.arange(start, start + n, dtype=np.uint64)
    # radius by inversion of the cumulative mass, redrawn while r > rmax
    r = np.empty(n, np.float32)
    todo = np.arange(n)
    attempt = 0
    while todo.size:
        u = uniform01(seed + 7919 * attempt, idx[todo], 0).astype(np.float64)
        u = np.clip(u, 1e-7, 1.0 - 1e-7)
        rr = scale / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        ok = rr <= rmax
        r[todo[ok]] = rr[ok].astype(np.float32)
        todo = todo[~ok]
        attempt += 1
    pos = np.asarray(centre, np.float32)[None, :] + r[:, None] * _isotropic(uniform01(seed, idx, 1), uniform01(seed, idx, 2))
    # speed: q = v / v_esc from g(q) = q^2 (1-q^2)^(7/2) by rejection (Aarseth, Henon, Wielen 1974)
    q = np.empty(n, np.float32)
    todo = np.arange(n)
    attempt = 0
    while todo.size:
        x = uniform01(seed + 104729 * (attempt + 1), idx[todo], 5).astype(np.float64)
        y = 0.1 * uniform01(seed + 104729 * (attempt + 1), idx[todo], 6).astype(np.float64)
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q[todo[ok]] = x[ok].astype(np.float32)
        todo = todo[~ok]
        attempt += 1
    vesc = np.sqrt(2.0 * total_mass) * (r.astype(np.float64) ** 2 + scale * scale) ** -0.25
    vel = (q * vesc).astype(np.float32)[:, None] * _isotropic(uniform01(seed, idx, 3), uniform01(seed, idx, 4))
    vel = vel + np.asarray(bulk_v, np.float32)[None, :]
    mass = np.full(n, total_mass / n_total, np.float32)
    return _pack(pos.astype(np.float32), vel.astype(np.float32), mass, mass)


def plummer(n, seed=42, start=0, n_total=None):
    """Plummer sphere: scale a = 1/32, truncated at r = 0.45,
    centred in the unit cube, equal masses 1/N (total mass 1, G = 1), virial
    isotropic velocities."""
    return _plummer(n, seed, start, (0.5, 0.5, 0.5), 1.0 / 32.0, 0.45, 1.0, (0, 0, 0), n_total or n)


def two_galaxies(n, seed=42, start=0, n_total=None):
    """Two Plummer spheres of N/2 at x = 0.3 / 0.7 approaching at +-0.05 (config 4). With n_total, particles
    [start, start + n) of the N = n_total set (ranks generate disjoint slices of one global set)."""
    N = n_total or n
    h = N // 2
    lo, hi = start, start + n
    na = max(0, min(hi, h) - lo)
    nb = n - na
    parts = []
    if na:
        parts.append(_plummer(na, seed, lo, (0.3, 0.5, 0.5), 1.0 / 32.0, 0.28, 0.5, (0.05, 0, 0), h))
    if nb:
        parts.append(_plummer(nb, seed + 1, max(lo, h), (0.7, 0.5, 0.5), 1.0 / 32.0, 0.28, 0.5, (-0.05, 0, 0), N - h))
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, PARTICLE_FLOATS), np.float32)


GENERATORS = {"uniform": uniform_cube, "plummer": plummer, "two_galaxies": two_galaxies}


def generate(kind, n_total, start=0, count=None):
    """Particles [start, start + count) of the n_total-particle workload `kind`: every rank builds its own slice."""
    count = n_total - start if count is None else count
    if kind == "uniform":
        return uniform_cube(count, start=start)
    return GENERATORS[kind](count, start=start, n_total=n_total)


def force_constant(kind, n):
    """Force constant that keeps each workload dynamically tame at dt = 1e-3.

    The reference demo (src/main.cpp:23-69: masses 1..10, unit force constant, dt = 0.001)
    is violently unstable for any sizeable N — the free-fall time of the cube is far
    below dt, so after one step every particle has left the box (and the reference has
    no boundary handling). Multi-step runs of the uniform cube therefore use
    G = 1 / (total mass) (free-fall time of order 1); single force evaluations are
    unaffected because G only scales the accelerations. The Plummer models are built in
    virial equilibrium for G = 1 and total mass 1."""
    if kind == "uniform":
        return 1.0 / (5.5 * n)
    return 1.0
