"""pytest plugin: imports the reference through the shim and replaces its FDM
differentiator / explicit integrators / FDM operator by the oracle adapters, so
that the reference's OWN test files exercise the oracle."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

import refshim  # noqa: E402

refshim.install()

import pararealml.operators.fdm as fdm_pkg  # noqa: E402
import pararealml.operators.fdm.fdm_operator as fo  # noqa: E402
import pararealml.operators.fdm.numerical_differentiator as nd  # noqa: E402
import pararealml.operators.fdm.numerical_integrator as ni  # noqa: E402

import oracle_adapter as oa  # noqa: E402

nd.ThreePointCentralDifferenceMethod = oa.OracleThreePointCentralDifferenceMethod
ni.ForwardEulerMethod = oa.OracleForwardEulerMethod
ni.ExplicitMidpointMethod = oa.OracleExplicitMidpointMethod
ni.RK4 = oa.OracleRK4
if os.environ.get("PML_ALIAS_FDM_OPERATOR") == "1":
    fo.FDMOperator = oa.OracleFDMOperatorAdapter
    fdm_pkg.FDMOperator = oa.OracleFDMOperatorAdapter
for name in ("ThreePointCentralDifferenceMethod",):
    setattr(fdm_pkg, name, getattr(nd, name))
for name in ("ForwardEulerMethod", "ExplicitMidpointMethod", "RK4"):
    setattr(fdm_pkg, name, getattr(ni, name))
