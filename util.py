"""Top-level drop-in for `from util import save_video, seed_everything`
(run_tokenflow_pnp.py:17, run_tokenflow_sdedit.py:16)."""
from tokenflow_b200.util import (isinstance_str, batch_cosine_sim, seed_everything, save_video, load_imgs,  # noqa: F401
                                 save_video_frames, add_dict_to_yaml_file)
