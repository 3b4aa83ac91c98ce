"""ctypes binding of the C ABI declared in include/scisim_b200.h.

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing the
host classes raises, so a CPU path can never silently stand in for the product.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscisim_b200.so")

SG_OK = 0
SG_ERR_INVALID, SG_ERR_CUDA, SG_ERR_UNSUPPORTED, SG_ERR_INTERNAL, SG_ERR_REBALANCE = 1, 2, 3, 4, 5
SG_MAP_NONE = -1
SG_MAP_SYMPLECTIC_EULER, SG_MAP_VERLET, SG_MAP_SPLIT_HAM, SG_MAP_DMV, SG_MAP_EXPONENTIAL_EULER = 0, 1, 2, 3, 4
SG_MAP_M_UPDATED = 0x100  # rigidbody3d flows after the first: M as updateMandMinv leaves it (include/scisim_b200.h)
SG_BALL_BALL, SG_BALL_DRUM, SG_BALL_PLANE = 0, 1, 2
SG_BALL_BALL_TELEPORTED, SG_BALL_BALL_KICK_TELEPORTED = 3, 4
SG_NO_PORTAL, SG_PLANE_B_BIT = 0xFFFFFFFF, 0x80000000
SG_CIRCLE_CIRCLE_TELEPORTED, SG_CIRCLE_CIRCLE_KICK_TELEPORTED = 25, 26
SG_SPHERE_SPHERE_TELEPORTED, SG_KINEMATIC_OBJECT_SPHERE_TELEPORTED = 19, 30
SG_SPHERE_SPHERE, SG_KINEMATIC_SPHERE_SPHERE, SG_BODY_BODY, SG_KINEMATIC_BODY_BODY, SG_PLANE_SPHERE, SG_PLANE_BOX, SG_PLANE_BODY = 10, 11, 12, 13, 14, 15, 16
SG_OUT_NORMALS, SG_OUT_POINTS, SG_OUT_DEPTHS, SG_OUT_CANDIDATES, SG_OUT_ALL = 1, 2, 4, 8, 15
SG_IN_RESIDENT = 256

c_dp = C.POINTER(C.c_double)
c_up = C.POINTER(C.c_uint32)


class SgPairs(C.Structure):
    _fields_ = [("n", C.c_uint64), ("ij", c_up)]


class SgContacts(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("n_candidates", C.c_uint64), ("n_active", C.c_uint64), ("n_body_body", C.c_uint64),
                ("n_drum", C.c_uint64), ("n_plane", C.c_uint64), ("type", c_up), ("i", c_up), ("j", c_up),
                ("n", c_dp), ("p", c_dp), ("depth", c_dp), ("cand_ij", c_up), ("aux", c_up)]


class SgAssembly(C.Structure):
    _fields_ = [("n_constraints", C.c_uint64), ("n_dofs", C.c_uint64), ("n_nnz", C.c_uint64), ("n_outer", C.POINTER(C.c_int32)), ("n_inner", C.POINTER(C.c_int32)), ("n_values", c_dp),
                ("q_nnz", C.c_uint64), ("q_outer", C.POINTER(C.c_int32)), ("q_inner", C.POINTER(C.c_int32)), ("q_values", c_dp), ("bases", c_dp)]


class SgTeleported(C.Structure):
    _fields_ = [("n_boxes", C.c_uint64), ("box_body", c_up), ("box_portal", c_up), ("n_regular", C.c_uint64), ("n_teleported", C.c_uint64),
                ("portal0", c_up), ("portal1", c_up), ("x0", c_dp), ("x1", c_dp), ("kick", c_dp), ("delta0", c_dp), ("delta1", c_dp)]


class SciSimB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Loads libscisim_b200.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SciSimB200Error("%s not built: run `python -m scisim_b200.build` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sigs = {
        "sg_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "sg_destroy": (None, [vp]),
        "sg_last_error": (C.c_char_p, [vp]),
        "sg_synchronize": (C.c_int, [vp]),
        "sg_host_alloc": (C.c_int, [vp, C.c_uint64, C.POINTER(vp)]),
        "sg_host_free": (C.c_int, [vp, vp]),
        "sg_stream": (vp, [vp]),
        "sg_profile_enable": (C.c_int, [vp, C.c_int]),
        "sg_profile_reset": (C.c_int, [vp]),
        "sg_profile_count": (C.c_int, [vp]),
        "sg_profile_get": (C.c_int, [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), c_dp, c_dp]),
        "sg_launch_count": (C.c_uint64, [vp]),
        "sg_timer_begin": (C.c_int, [vp]),
        "sg_timer_end": (C.c_int, [vp, c_dp]),
        "sg_flush_l2": (C.c_int, [vp]),
        "sg_candidate_pairs": (C.c_int, [vp, C.c_int, C.c_uint32, vp, C.POINTER(SgPairs)]),
        "sg_ball2d_set_bodies": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_ball2d_set_gravity": (C.c_int, [vp, vp]),
        "sg_ball2d_set_planes": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_ball2d_set_drums": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_ball2d_set_portals": (C.c_int, [vp, C.c_uint32, vp, vp, vp, vp, vp, vp]),
        "sg_ball2d_update_portals": (C.c_int, [vp, C.c_double, vp]),
        "sg_ball2d_enforce_portals": (C.c_int, [vp, vp, vp]),
        "sg_ball2d_teleported": (C.c_int, [vp, C.POINTER(SgTeleported)]),
        "sg_ball2d_flow": (C.c_int, [vp, C.c_int, vp, vp, C.c_double, vp, vp]),
        "sg_ball2d_active_set": (C.c_int, [vp, vp, vp, C.c_uint32, C.POINTER(SgContacts)]),
        "sg_ball2d_upload": (C.c_int, [vp, vp, vp]),
        "sg_ball2d_step": (C.c_int, [vp, C.c_int, C.c_double, C.POINTER(SgContacts)]),
        "sg_ball2d_fetch": (C.c_int, [vp, C.c_uint32, vp, vp, C.POINTER(SgContacts)]),
        "sg_ball2d_slab_init": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]),
        "sg_ball2d_slab_set_gids": (C.c_int, [vp, vp, vp]),
        "sg_ball2d_slab_upload_q1": (C.c_int, [vp, vp]),
        "sg_ball2d_slab_stats": (C.c_int, [vp, vp]),
        "sg_ball2d_assemble": (C.c_int, [vp, C.c_uint32, C.POINTER(SgAssembly)]),
        "sg_ball2d_cache_store": (C.c_int, [vp, C.c_uint32, vp]),
        "sg_ball2d_cache_lookup": (C.c_int, [vp, C.c_uint32, vp, C.POINTER(C.c_uint64)]),
        "sg_ball2d_cache_clear": (C.c_int, [vp]),
        "sg_ball2d_state_serialize": (C.c_int, [vp, C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint64)]),
        "sg_ball2d_state_deserialize": (C.c_int, [vp, vp, C.c_uint64]),
        "sg_rb3d_state_serialize": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint64)]),
        "sg_rb3d_state_deserialize": (C.c_int, [vp, vp, C.c_uint64]),
        "sg_rb3d_set_mesh_snapshot": (C.c_int, [vp, C.c_uint32, vp, C.c_uint64]),
        "sg_rb2d_state_serialize": (C.c_int, [vp, C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint64)]),
        "sg_rb2d_state_deserialize": (C.c_int, [vp, vp, C.c_uint64]),
        "sg_ball2d_fetch_state": (C.c_int, [vp, vp, vp]),
        "sg_slab_partition": (C.c_int, [C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, vp]),
        "sg_slab_limits": (C.c_int, [C.c_uint32, vp, C.c_uint32, vp]),
        "sg_slab_merge_dest": (C.c_int, [C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp]),
        "sg_slab_merge_static_dest": (C.c_int, [C.c_uint32, vp, vp, vp, vp, vp]),
        "sg_create_multi": (C.c_int, [C.POINTER(vp), C.c_int, vp]),
        "sg_destroy_multi": (None, [vp]),
        "sg_multi_last_error": (C.c_char_p, [vp]),
        "sg_multi_n_gpus": (C.c_int, [vp]),
        "sg_multi_context": (vp, [vp, C.c_int]),
        "sg_multi_set_rebalance": (C.c_int, [vp, C.c_uint32, C.c_uint32]),
        "sg_multi_partition_info": (C.c_int, [vp, vp, vp, vp, vp]),
        "sg_multi_ball2d_set_bodies": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_multi_ball2d_set_gravity": (C.c_int, [vp, vp]),
        "sg_multi_ball2d_set_planes": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_multi_ball2d_set_drums": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_multi_ball2d_flow": (C.c_int, [vp, C.c_int, vp, vp, C.c_double, vp, vp]),
        "sg_multi_ball2d_active_set": (C.c_int, [vp, vp, vp, C.c_uint32, C.POINTER(SgContacts)]),
        "sg_multi_ball2d_upload": (C.c_int, [vp, vp, vp]),
        "sg_multi_ball2d_step": (C.c_int, [vp, C.c_int, C.c_double, C.POINTER(SgContacts)]),
        "sg_multi_ball2d_fetch": (C.c_int, [vp, C.c_uint32, vp, vp, C.POINTER(SgContacts)]),
        "sg_ball2d_slab_flow": (C.c_int, [vp, C.c_int, C.c_double, vp]),
        "sg_ball2d_slab_pack": (C.c_int, [vp, vp, vp, C.c_uint32, vp]),
        "sg_ball2d_slab_unpack": (C.c_int, [vp, C.c_int, vp]),
        "sg_ball2d_slab_mailbox": (C.c_int, [vp, C.POINTER(C.c_void_p), vp]),
        "sg_ball2d_slab_connect": (C.c_int, [vp, C.c_int, vp, vp, C.c_int]),
        "sg_ball2d_slab_exchange": (C.c_int, [vp, C.c_int]),
        "sg_ball2d_slab_disconnect": (C.c_int, [vp]),
        "sg_ball2d_slab_detect": (C.c_int, [vp, C.POINTER(SgContacts), vp]),
        "sg_rb2d_set_geometry": (C.c_int, [vp, C.c_uint32, vp, vp, vp]),
        "sg_rb2d_set_bodies": (C.c_int, [vp, C.c_uint32, vp, vp, vp]),
        "sg_rb2d_set_gravity": (C.c_int, [vp, vp]),
        "sg_rb2d_set_planes": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_rb2d_set_portals": (C.c_int, [vp, C.c_uint32, vp, vp, vp, vp, vp, vp]),
        "sg_rb2d_update_portals": (C.c_int, [vp, C.c_double, vp]),
        "sg_rb2d_enforce_portals": (C.c_int, [vp, vp, vp]),
        "sg_rb2d_teleported": (C.c_int, [vp, C.POINTER(SgTeleported)]),
        "sg_rb2d_flow": (C.c_int, [vp, C.c_int, vp, vp, C.c_double, vp, vp]),
        "sg_rb2d_active_set": (C.c_int, [vp, vp, vp, C.c_uint32, C.POINTER(SgContacts)]),
        "sg_rb2d_upload": (C.c_int, [vp, vp, vp]),
        "sg_rb2d_step": (C.c_int, [vp, C.c_int, C.c_double, C.POINTER(SgContacts)]),
        "sg_rb2d_fetch": (C.c_int, [vp, C.c_uint32, vp, vp, C.POINTER(SgContacts)]),
        "sg_rb3d_set_geometry": (C.c_int, [vp, C.c_uint32, vp, vp, vp, vp]),
        "sg_rb3d_add_mesh": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp, vp, C.POINTER(C.c_uint32)]),
        "sg_rb3d_set_bodies": (C.c_int, [vp, C.c_uint32, vp, vp, vp, vp]),
        "sg_rb3d_set_gravity": (C.c_int, [vp, vp]),
        "sg_rb3d_set_planes": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "sg_rb3d_set_portals": (C.c_int, [vp, C.c_uint32, vp, vp, vp, vp, vp]),
        "sg_rb3d_enforce_portals": (C.c_int, [vp, vp]),
        "sg_rb3d_teleported": (C.c_int, [vp, C.POINTER(SgTeleported)]),
        "sg_rb3d_update_m_and_minv": (C.c_int, [vp, vp, vp, vp]),
        "sg_rb3d_flow": (C.c_int, [vp, C.c_int, vp, vp, C.c_double, vp, vp]),
        "sg_rb3d_active_set": (C.c_int, [vp, vp, vp, C.c_uint32, C.POINTER(SgContacts)]),
        "sg_rb3d_upload": (C.c_int, [vp, vp, vp]),
        "sg_rb3d_step": (C.c_int, [vp, C.c_int, C.c_double, C.POINTER(SgContacts)]),
        "sg_rb3d_fetch": (C.c_int, [vp, C.c_uint32, vp, vp, C.POINTER(SgContacts)]),
        "sg_rb3d_slab_init": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]),
        "sg_rb3d_slab_mailbox": (C.c_int, [vp, C.POINTER(C.c_void_p), vp]),
        "sg_rb3d_slab_connect": (C.c_int, [vp, C.c_int, vp, vp, C.c_int]),
        "sg_rb3d_slab_disconnect": (C.c_int, [vp]),
        "sg_rb3d_slab_flow": (C.c_int, [vp, C.c_int, C.c_double]),
        "sg_rb3d_slab_exchange": (C.c_int, [vp, C.c_int]),
        "sg_rb3d_slab_detect": (C.c_int, [vp, C.POINTER(SgContacts), vp]),
        "sg_rb3d_set_cylinders": (C.c_int, [vp, C.c_uint32, vp, vp, vp]),
        "sg_rb3d_mesh_stats": (C.c_int, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    """Every function name declared in include/scisim_b200.h (used by the ABI test)."""
    import re
    hdr = open(os.path.join(HERE, "..", "include", "scisim_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", hdr)))
