"""meshode_b200 -- sm_100a (B200) implementation of MeshODE's data-parallel hot path:
distance-field build, trilinear distance/gradient lookup and edge-rigidity losses, behind
the reference's ``pyDeform`` API.  GPU only: there is no CPU fallback."""
__version__ = "0.1.0"
