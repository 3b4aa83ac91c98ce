"""CPU checks of the reasoning behind three kernel shortcuts (formulas restated in numpy / plain Python and run on adversarial inputs; what the kernels do with them is checked on the GPU against the oracle).

(1) pass 1 far-side pruning (scisim_b200/csrc/sg_broadphase.cuh, sg_bp_count_l1): bodies are binned by the
LOWER corner of their box, cell = uint32( ( lo - origin ) / h ) clamped, h >= every extent; a body drops the next cell column (row) from its walk when
its upper bound, rounded UP to float, lies below  origin + ( c + 1 ) h  minus a margin.  Restated here in numpy with the kernel's expressions and run
on adversarial inputs -- boxes that end exactly on, one ulp below and one ulp above cell edges, huge offsets, tiny cells: no overlapping pair may fall
outside the pruned ranges of either of its bodies.  (What the kernel does with these formulas is checked on the GPU against the oracle; this test is
about the formulas.)"""
import numpy as np


def _cells(lo, origin, h, dims):
    v = np.floor((lo - origin) / h)          # lo >= origin, so the unsigned cast is a floor
    v = np.minimum(v, dims - 1)
    return v.astype(np.int64)


def _pruned_hi_cell(c, hi, origin, h, dims):
    """last cell column the walk of a body in column c visits: c + 1 unless the box ends below that column's lower edge"""
    hi_f = np.nextafter(hi.astype(np.float32), np.float32(np.inf), where=hi.astype(np.float32).astype(np.float64) < hi, out=hi.astype(np.float32).copy()).astype(np.float64)  # __double2float_ru
    edge = (c + 1).astype(np.float64) * h
    skip = hi_f < (origin + edge) - 1.0e-9 * (np.abs(origin) + edge)
    last = np.where(skip, c, np.minimum(c + 1, dims - 1))
    return last


def _check(lo, ext, h):
    hi = lo + ext
    origin = lo.min()
    assert np.all(ext <= h)
    dims = int(np.floor((lo.max() - origin) / h)) + 1
    c = _cells(lo, origin, h, dims)
    last = _pruned_hi_cell(c, hi, origin, h, dims)
    first = np.maximum(c - 1, 0)
    # every overlapping pair (closed intervals, as AABB::overlaps)
    order = np.argsort(lo)
    lo_s, hi_s, c_s, f_s, l_s = lo[order], hi[order], c[order], first[order], last[order]
    n = lo.shape[0]
    bad = 0
    for i in range(n):
        j = i + 1
        while j < n and lo_s[j] <= hi_s[i]:
            # i and j overlap ( lo_j <= hi_i and lo_i <= lo_j <= hi_j ): each must find the other's cell inside its own range
            if not (f_s[i] <= c_s[j] <= l_s[i]) or not (f_s[j] <= c_s[i] <= l_s[j]):
                bad += 1
            j += 1
    return bad, int((last == c).sum())


def test_far_side_pruning_never_hides_an_overlapping_box():
    rng = np.random.default_rng(12)
    total_pruned = 0
    for trial in range(60):
        n = 400
        scale = 10.0 ** rng.integers(-3, 7)           # coordinates from 1e-3 to 1e6
        offset = rng.choice([0.0, 1.0e3, -7.7e5, 3.0e9]) * (1.0 if trial % 3 else 0.0)
        h_true = scale * rng.uniform(0.5, 2.0)
        ext = h_true * rng.uniform(0.05, 1.0, n)
        ext[rng.integers(0, n)] = h_true              # the largest extent defines the cell size, as in sg_layout_grid
        h = ext.max() * (1.0 + 9.5367431640625e-07)
        lo = offset + rng.uniform(0.0, 40.0 * h, n)
        # adversarial: upper bounds exactly on, just below and just above cell edges (origin = lo.min() is only known afterwards: anchor one body at it)
        lo[0] = lo.min() - 0.25 * h
        k = rng.integers(1, 38, n // 2)
        edge = lo[0] + k * h
        sel = np.arange(1, n // 2 + 1)
        nudge = rng.integers(-2, 3, n // 2)
        hi_target = edge.copy()
        for s in range(n // 2):
            for _ in range(abs(int(nudge[s]))):
                hi_target[s] = np.nextafter(hi_target[s], np.inf if nudge[s] > 0 else -np.inf)
        lo[sel] = hi_target - ext[sel]
        lo[sel] = np.maximum(lo[sel], lo[0])
        bad, pruned = _check(lo, ext, h)
        assert bad == 0, (trial, bad)
        total_pruned += pruned
    assert total_pruned > 2000   # the pruning does happen on these inputs


def _round_up_f32(x):
    f = x.astype(np.float32)
    return np.where(f.astype(np.float64) < x, np.nextafter(f, np.float32(np.inf)), f).astype(np.float32)


def _round_down_f32(x):
    f = x.astype(np.float32)
    return np.where(f.astype(np.float64) > x, np.nextafter(f, np.float32(-np.inf)), f).astype(np.float32)


def test_float_gap_certainty_implies_fp64_overlap():
    """sg_box_gap_certain: with boxes rounded outward to floats ( lo_f <= lo, hi <= hi_f ), a float gap hi_f - lo_f above 4e-7 ( |hi_f| + |lo_f| ) + 1e-37
    must imply hi >= lo for the FP64 values -- the condition under which pass 1 skips rebuilding the FP64 boxes.  Swept over magnitudes from
    subnormal floats to 1e30 with hi placed within twenty float spacings of lo on either side."""
    rng = np.random.default_rng(3)
    n = 400_000
    mag = 10.0 ** rng.uniform(-44, 30, n)
    lo = mag * rng.choice([-1.0, 1.0], n) * rng.uniform(0.5, 1.0, n)
    ulp32 = np.maximum(np.abs(lo) * 2.0 ** -23, 2.0 ** -149)
    hi = lo + ulp32 * rng.uniform(-20.0, 20.0, n)
    hi_f, lo_f = _round_up_f32(hi), _round_down_f32(lo)
    with np.errstate(over="ignore", invalid="ignore"):
        gap = (hi_f - lo_f).astype(np.float32)
        rhs = (np.float32(4.0e-7) * (np.abs(hi_f) + np.abs(lo_f)).astype(np.float32)).astype(np.float32) + np.float32(1.0e-37)
        certain = gap > rhs
    assert certain.sum() > n // 20 and (~certain).sum() > n // 20
    assert np.all(hi[certain] >= lo[certain])
    # and it is not vacuous: among the undecided ones both outcomes occur
    und = ~certain
    assert (hi[und] >= lo[und]).any() and (hi[und] < lo[und]).any()


def _angle_axis_matrix(sn, cs, axis):
    """Eigen 3.3.4 AngleAxis::toRotationMatrix, expression for expression (Eigen/src/Geometry/AngleAxis.h), with sin / cos given"""
    ax, ay, az = axis
    sx, sy, sz = sn * ax, sn * ay, sn * az
    c1x, c1y, c1z = (1.0 - cs) * ax, (1.0 - cs) * ay, (1.0 - cs) * az
    m = [[0.0] * 3 for _ in range(3)]
    tmp = c1x * ay
    m[0][1] = tmp - sz; m[1][0] = tmp + sz
    tmp = c1x * az
    m[0][2] = tmp + sy; m[2][0] = tmp - sy
    tmp = c1y * az
    m[1][2] = tmp - sx; m[2][1] = tmp + sx
    m[0][0] = c1x * ax + cs; m[1][1] = c1y * ay + cs; m[2][2] = c1z * az + cs
    return m


def test_structured_splitham_rotations_equal_the_full_products():
    """splitham_axis_step (scisim_b200/csrc/sg_rb3d.cu): SplitHamMap's rotations are about -e_z, -e_y, -e_x; the kernel writes the products
    R <- R * AngleAxis( -a, -e_K ).matrix() and p <- AngleAxis( a, -e_K ).matrix() * p in structured form.  With the same sin / cos, the structured
    form must equal the full 3x3 products of Eigen's matrix -- every entry, bit for bit (signs of exact zeros aside)."""
    import math
    rng = np.random.default_rng(9)
    for trial in range(3000):
        R = rng.normal(size=(3, 3)).tolist()
        p = rng.normal(size=3).tolist()
        a = float(rng.uniform(-3.5, 3.5)) if trial % 10 else 0.0
        sn, cs = math.sin(a), math.cos(a)
        e = (1.0 - cs) + cs
        for K in range(3):
            axis = [-0.0, -0.0, -0.0]
            axis[K] = -1.0
            B = _angle_axis_matrix(-sn, cs, axis)     # AngleAxis( -a, axis ): sin( -a ) = -sin( a )
            A = _angle_axis_matrix(sn, cs, axis)
            full_R = [[(R[r][0] * B[0][c] + R[r][1] * B[1][c]) + R[r][2] * B[2][c] for c in range(3)] for r in range(3)]
            full_p = [(A[r][0] * p[0] + A[r][1] * p[1]) + A[r][2] * p[2] for r in range(3)]
            S = [[0.0] * 3 for _ in range(3)]
            for r in range(3):
                a0, a1, a2 = R[r]
                if K == 2:
                    S[r] = [a0 * cs + a1 * sn, a0 * (-sn) + a1 * cs, a2 * e]
                elif K == 1:
                    S[r] = [a0 * cs + a2 * (-sn), a1 * e, a0 * sn + a2 * cs]
                else:
                    S[r] = [a0 * e, a1 * cs + a2 * sn, a1 * (-sn) + a2 * cs]
            px, py, pz = p
            if K == 2:
                sp = [cs * px + sn * py, (-sn) * px + cs * py, e * pz]
            elif K == 1:
                sp = [cs * px + (-sn) * pz, e * py, sn * px + cs * pz]
            else:
                sp = [e * px, cs * py + sn * pz, (-sn) * py + cs * pz]
            assert np.array_equal(np.array(S), np.array(full_R)), (trial, K)
            assert np.array_equal(np.array(sp), np.array(full_p)), (trial, K)
