"""learning_to_adapt_b200 -- B200-native (sm_100a) MPC planning engine behind the learning_to_adapt API.

Only the planner hot path lives here: MPCController.get_action(s) (random shooting / CEM), (Meta)MLPDynamicsModel
.predict / .adapt / .switch_to_pre_adapt, as hand-written CUDA kernels behind a C ABI (include/l2a_b200.h).
"""
__version__ = "0.1.0"
