"""Input formats of the hot path (reference ``dataio/``): raw camera streams -> frame batches."""
from .octopus import OctopusData, write_octopus_stream   # noqa: F401
from .micromanager import (MicromanagerMetadataParser, MicromanagerReader, read_tiff, write_tiff,   # noqa: F401
                           write_micromanager_position)
