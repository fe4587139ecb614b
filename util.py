"""Top-level drop-in for `from util import save_video, seed_everything`
(run_tokenflow_pnp.py:17, run_tokenflow_sdedit.py:16)."""
from tokenflow_b200.util import isinstance_str, batch_cosine_sim, seed_everything, save_video  # noqa: F401
