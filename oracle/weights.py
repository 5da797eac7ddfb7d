"""Re-export of hirest_b200.synthetic (seeded synthetic weights / inputs) for the oracle-side scripts."""
from hirest_b200.synthetic import *  # noqa: F401,F403
from hirest_b200.synthetic import EVA_G14, EVA_TINY, make_eva_state_dict, make_frames, make_text_state_dict, \
    make_tokens, make_visual_state_dict  # noqa: F401
