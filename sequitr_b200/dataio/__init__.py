"""Input formats of the hot path (reference ``dataio/``): raw camera streams -> frame batches."""
from .octopus import OctopusData, write_octopus_stream   # noqa: F401
