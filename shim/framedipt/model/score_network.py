"""framedipt.model.score_network of the reference (framedipt/model/score_network.py), served by the B200 path."""
from framedipt_b200.runtime import index_embedding as get_index_embedding  # noqa: F401
from framedipt_b200.runtime import timestep_embedding as get_timestep_embedding  # noqa: F401
from framedipt_b200.score_network import ScoreNetwork  # noqa: F401
