"""TEST INFRASTRUCTURE — re-export of the seeded synthetic-input helpers (they live in controlar_b200/synthetic.py so that
bench.py's product arm does not import anything from oracle/)."""
from controlar_b200.synthetic import *  # noqa: F401,F403
from controlar_b200.synthetic import text_inputs, class_inputs, control_map, xl_ctrl_in, train_attn_mask, code_inputs  # noqa: F401
