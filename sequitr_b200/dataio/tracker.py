"""Tracker hand-off formats (reference ``dataio/tracker.py``): the
containers a Sequitr job reads back from BayesianTracker once the centroid tables written by
``utils.CentroidWriter`` have been linked into trajectories.  Host-side only -- nothing here touches
the GPU; it closes the loop  frames -> masks -> centroids -> (btrack) -> tracks  for users of the
hot path.

Same names and behaviour as the reference: ``FATE_LABELS`` (:32-33), ``Track`` (:36-127),
``read_JSON`` (:133-181), ``read_XML`` (:186-222), ``write_XML`` (:229-260).  Differences, all bug
fixes of Python-2-only or unreachable code paths:
* ``read_XML``: the reference stores the ``<n>`` list with ``setattr(track, 'n', ...)``, which fails
  on the read-only ``n`` property (:60) and is swallowed by a bare ``except`` (:213-218), so frame
  numbers written by ``write_XML`` never come back; here ``<n>`` fills ``Track.t``.
* string checks use ``str`` (``basestring`` does not exist under Python 3); nothing is printed.
"""
import ast
import json
import os
import zipfile
import xml.etree.ElementTree as ET

FATE_LABELS = ['null', 'initializing', 'terminating', 'link',
               'mitosis', 'apoptosis', 'dead', 'merging', 'undefined']

_COPIED = ('ID', 'length', 'parent', 'fate', 'cell_type', 'children')
_PER_FRAME = ('x', 'y', 'z', 't', 'label')


class Track(object):
    """ Dummy track object: a container for the tracker's output (reference :36-50). """

    def __init__(self, ID=None):
        self.ID = ID
        self.x = None
        self.y = None
        self.z = None
        self.t = None
        self.length = None
        self.label = None
        self.parent = None
        self.children = []
        self.fate = None
        self.cell_type = None
        self.neighborhood = []
        self.filename = None

    @property
    def n(self):
        return self.t

    def in_frame(self, frame):
        """ whether this track is present in a certain frame of the movie """
        return frame in self.t

    def __len__(self):
        return self.length

    @property
    def fate_as_string(self):
        return FATE_LABELS[self.fate] if self.fate else 'undefined'

    def get_copy_at_frame(self, frame):
        """ the cell position and state at one frame, as a new Track whose ``ref`` is this one """
        if not self.in_frame(frame):
            return None
        idx = list(self.n).index(frame)
        T = Track()
        T.ref = self
        for p in _COPIED:
            setattr(T, p, getattr(self, p))
        for p in _PER_FRAME:
            param = getattr(self, p)
            if param is not None:
                setattr(T, p, param[idx])
        return T

    def get_neighborhood_attr(self, attr):
        if not isinstance(attr, str):
            raise TypeError('Attribute must be a string')
        return [n[attr] for n in self.neighborhood]

    def __getitem__(self, attr):
        """ Get an item by name! """
        if attr in self.__dict__:
            return getattr(self, attr)
        if self.neighborhood and attr in self.neighborhood[0]:
            return self.get_neighborhood_attr(attr)
        return None

    @staticmethod
    def from_dict(params):
        T = Track()
        for k, v in params.items():
            if k == 'n':                  # read-only alias of t
                k = 't'
            setattr(T, k, v)
        return T


def read_JSON(folder, cell_type):
    """ read tracks in JSON format

        tracks_<cell_type>.json is organized as follows:

        "<cell_type>":
            "files": [track_1_<cell_type>.json], ...],
            "path": <path>
            "zipped": bool  (the track files then live in tracks_<cell_type>.zip)

        Each track.json file is the actual track data, which is inserted into a Track object.
    """
    file_stats_fn = os.path.join(folder, "tracks_{}.json".format(cell_type))
    if not os.path.exists(file_stats_fn):
        raise IOError('Tracking data file not found: {}'.format(file_stats_fn))
    with open(file_stats_fn, 'r') as json_file:
        track_files = json.load(json_file)

    def _make(d, track_fn):
        d['cell_type'] = cell_type
        d['filename'] = track_fn
        return Track.from_dict(d)

    entry = track_files[cell_type]
    tracks = []
    if entry['zipped']:
        zip_fn = os.path.join(folder, "tracks_{}.zip".format(cell_type))
        with zipfile.ZipFile(zip_fn, 'r') as zipped_tracks:
            for track_fn in entry['files']:
                tracks.append(_make(json.loads(zipped_tracks.read(track_fn)), track_fn))
        return tracks
    for track_fn in entry['files']:
        with open(os.path.join(folder, track_fn), 'r') as track_file:
            tracks.append(_make(json.load(track_file), track_fn))
    return tracks


def read_XML(filename, cell_type=None):
    """ Load tracks from a sequitr XML file """
    if filename is None:
        return []
    if not isinstance(filename, str):
        raise TypeError("Filename must be specified as a string")
    if not filename.endswith((".xml", ".XML")):
        raise IOError("Tracking data must be in XML format")
    if not os.path.exists(filename):
        return []
    tracks = []
    for track in ET.parse(filename).getroot().findall('trajectory'):
        new_track = Track(ID=int(track.get('id')))
        new_track.cell_type = cell_type
        for prop in track:
            try:
                value = ast.literal_eval(prop.text)
            except (ValueError, SyntaxError):
                continue
            tag = {'class': 'label', 'n': 't'}.get(prop.tag, prop.tag)
            setattr(new_track, tag, value)
        tracks.append(new_track)
    return tracks


def write_XML(filename, tracks):
    """ write out the tracks to a new XML file (coordinates to one decimal, reference :243-244) """
    root = ET.Element("data", name=filename)
    for trk in tracks:
        if len(trk) < 1:
            continue
        txml = ET.SubElement(root, "trajectory", id=str(int(trk.ID)))
        ET.SubElement(txml, "length").text = str(len(trk))
        ET.SubElement(txml, "fate").text = str(trk.fate)
        ET.SubElement(txml, "x").text = str([float("{0:2.1f}".format(x)) for x in trk.x])
        ET.SubElement(txml, "y").text = str([float("{0:2.1f}".format(y)) for y in trk.y])
        ET.SubElement(txml, "n").text = str([int(t) for t in trk.n])
        ET.SubElement(txml, "class").text = str([l for l in trk.label])
        ET.SubElement(txml, "parent").text = str(trk.parent)
        ET.SubElement(txml, "children").text = str(trk.children)
        if trk.neighborhood:
            ET.SubElement(txml, "n_total").text = str([n for n in trk['n_total']])
            ET.SubElement(txml, "n_winner").text = str([n for n in trk['n_winner']])
            ET.SubElement(txml, "n_loser").text = str([n for n in trk['n_loser']])
            ET.SubElement(txml, "local_density").text = str(
                [float("{0:2.5f}".format(d)) for d in trk['local_density']])
            ET.SubElement(txml, "neighbors").text = str([[t.ID for t in refs] for refs in trk['refs']])
    ET.ElementTree(root).write(filename)
