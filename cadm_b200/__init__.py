"""cadm_b200 -- B200-native CEM/MPC planning engine for CaDM's planner hot path.

Public surface (mirrors the reference for this path):
    cadm_b200.dynamics.mlp_ensemble_cem_dynamics.MLPEnsembleCEMDynamicsModel        (PE-TS / vanilla)
    cadm_b200.dynamics.mlp_cadm_ensemble_cem_dynamics.MLPEnsembleCEMDynamicsModel   (CaDM)
    cadm_b200.policies.mpc_controller.MPCController
    cadm_b200.engine.PlannerEngine / PlannerConfig                                  (device-tensor API over the C ABI)
"""
__version__ = "0.1.0"
