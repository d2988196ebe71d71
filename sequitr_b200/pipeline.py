"""Drop-in for the image pipes of the reference ``sequitr/pipeline.py`` that sit on the hot path:
``ImagePipe`` (:162-189), ``ImagePipeline`` (:42-98) with its JSON round trip
(``save_image_pipeline`` :104-133, ``load_image_pipeline`` :137-154), the pre-inference pipes
``ImageOutliers`` / ``ImageNorm`` / ``ImageBGSubtract`` (GPU), ``ImageWeightMap`` (:455-479, GPU),
``ImageWeightMap2`` (:482-571, host) and the index-only pipes ``ImageFlip`` (:226-240) and
``ImageSample`` (:408-452).  ``ImageResize`` / ``ImageRotate`` / ``ImageBlur`` (scikit-image /
SciPy augmentation filters) are not part of this path; ``networks.unet.tr_augment`` is the GPU
augmentation.

``ImageWeightMapUNet`` is the north-star ``w_c + w0*exp(-(d1+d2)^2/2 sigma^2)`` map
on instance labels (GPU); ``weightmap.create_weightmaps(method='unet')`` selects it (the
default there stays the reference's ``ImageWeightMap2``).
"""
import inspect
import json
import sys
from collections import OrderedDict

import numpy as np

from . import ops


class ImagePipeline(object):
    """ ImagePipeline: chain ImagePipe objects (pipeline.py:42-98), save / load the chain's
    parameters as JSON. """

    def __init__(self, pipeline=[]):
        self.pipeline = pipeline

    @property
    def pipeline(self):
        return self._pipeline

    @pipeline.setter
    def pipeline(self, pipeline):
        if not isinstance(pipeline, list):
            pipeline = [pipeline]
        if any([not isinstance(p, ImagePipe) for p in pipeline]):
            raise TypeError('Pipeline contains non pipe objects')
        self._pipeline = pipeline

    def __call__(self, image):
        for pipe in self.pipeline:
            image = pipe(image)
        return image

    def __len__(self):
        return int(np.prod([len(p) for p in self.pipeline]))

    def update(self):
        for pipe in self.pipeline:
            pipe.update()

    def save(self, filename):
        """ Write out the pipeline as a JSON file """
        save_image_pipeline(filename, self)

    @staticmethod
    def load(filename):
        """ Read in the pipeline from a JSON file """
        return load_image_pipeline(filename)


def save_image_pipeline(filename, pipeline_object):
    """ Save out the parameters of an ImagePipeline as a JSON file (reference :104-133): one
    entry per pipe, class name -> constructor arguments read back from the attributes of the same
    name (like the reference, pipes of the same class overwrite each other). """
    if not isinstance(pipeline_object, ImagePipeline):
        raise TypeError('Pipeline must be of type ImagePipeline')
    if not filename.endswith('.json'):
        filename += '.json'
    pipes = []
    for pipe in pipeline_object.pipeline:
        pipe_args = [a for a in inspect.signature(pipe.__init__).parameters]
        vals = {}
        for a in pipe_args:
            v = getattr(pipe, a)
            vals[a] = list(v) if isinstance(v, tuple) else v
        pipes.append((pipe.__class__.__name__, vals))
    with open(filename, 'w') as json_file:
        json.dump({'ImagePipeline': OrderedDict(pipes)}, json_file, indent=2, separators=(',', ': '))


def load_image_pipeline(filename):
    """ Load and create an ImagePipeline object from a file (reference :137-154). """
    with open(filename, 'r') as json_file:
        pipes = json.load(json_file, object_pairs_hook=OrderedDict)
    pipeline = []
    for name, kwargs in pipes['ImagePipeline'].items():
        Pipe = getattr(sys.modules[__name__], name)
        if not (isinstance(Pipe, type) and issubclass(Pipe, ImagePipe)):
            raise TypeError('%s is not an image pipe' % name)
        pipeline.append(Pipe(**kwargs))
    return ImagePipeline(pipeline)


class ImagePipe(object):
    """ ImagePipe: primitive image pipe (pipeline.py:162-189). """

    def __init__(self):
        self.iter = 0

    def __call__(self, image):
        image = np.asarray(image)
        if image.ndim < 3:
            image = image[..., np.newaxis].astype('float32')
        return self.pipe(image)

    def pipe(self, image):
        raise NotImplementedError('Image pipe is not defined.')

    def __len__(self):
        return 1

    def update(self):
        self.iter = (self.iter + 1) % len(self)


class ImageFlip(ImagePipe):
    """ ImageFlip: mirror flips in sequence, for data augmentation (reference :226-240). """

    def __init__(self):
        ImagePipe.__init__(self)
        self.flips = [[], [np.fliplr], [np.flipud], [np.fliplr, np.flipud]]

    def pipe(self, image):
        for flip in self.flips[self.iter]:
            image = flip(image)
        return image

    def __len__(self):
        return len(self.flips)


class ImageSample(ImagePipe):
    """ ImageSample: randomly placed regions of interest; the positions are kept so that the same
    regions of corresponding labels or weights can be taken (reference :408-452). """

    def __init__(self, samples=16, ROI_size=(512, 512)):
        ImagePipe.__init__(self)
        self.samples = samples
        self.ROI_size = ROI_size
        self.im_size = None
        self.boundary = int(ROI_size[0] / 2.)
        self.coords = None

    def pipe(self, image):
        sampled = np.zeros((self.samples, self.ROI_size[0], self.ROI_size[1], image.shape[-1]))
        self.im_size = image.shape
        if not self.coords:
            self.update()
        b = self.boundary
        for sample, (x, y) in enumerate(self.coords):
            sampled[sample, ...] = image[x - b:x + b, y - b:y + b, ...]
        return sampled

    def update(self):
        b = self.boundary
        x = np.random.randint(b, high=self.im_size[0] - b, size=(self.samples,))
        y = np.random.randint(b, high=self.im_size[1] - b, size=(self.samples,))
        self.coords = list(zip(x, y))

    def __len__(self):
        return self.samples


def _binary_plane(image, who):
    if image.ndim != 3 or image.shape[-1] != 1:
        raise NotImplementedError('%s: single-channel (H,W) / (H,W,1) masks only' % who)
    m = image[..., 0]
    if not np.all((m == 0) | (m == 1)):
        raise ValueError('%s: a {0,1} mask is required (weightmap.py:203 passes a bool image)' % who)
    return np.ascontiguousarray(m != 0).astype(np.uint8)


class ImageOutliers(ImagePipe):
    """ ImageOutliers (pipeline.py:266-296): remove hot pixels by comparing the raw image with a
    ``sigma`` x ``sigma`` median-filtered copy; where they differ by more than ``threshold`` the
    pixel takes the median value.  Runs on the GPU (``sq_image_outliers``), bit-exact with the
    reference's SciPy path; like the reference it updates ``image`` in place and returns it. """

    def __init__(self, sigma=2, threshold=5.):
        ImagePipe.__init__(self)
        self.sigma = sigma
        self.threshold = threshold

    def pipe(self, image):
        image[...] = ops.image_pipe_host('outliers', image, size=self.sigma, threshold=self.threshold)
        return image


class ImageNorm(ImagePipe):
    """ ImageNorm (pipeline.py:338-356): subtract the mean and divide by the standard deviation, per
    channel.  Runs on the GPU (``sq_image_norm``); in place like the reference. """

    def __init__(self):
        ImagePipe.__init__(self)
        self.epsilon = 1e-99

    def pipe(self, image):
        image[...] = ops.image_pipe_host('norm', image)
        return image


class ImageBGSubtract(ImagePipe):
    """ ImageBGSubtract (pipeline.py:360-405): estimate the background as a second-order polynomial
    surface (least squares over every pixel) and subtract it.  Runs on the GPU
    (``sq_image_bgsubtract``); returns a new (H,W,1) float64 array like the reference. """

    def __init__(self):
        ImagePipe.__init__(self)

    def pipe(self, image):
        if image.shape[-1] != 1:
            # np.ravel(image) against an (H*W)-row design matrix (pipeline.py:398) only works for 1 channel
            raise ValueError('ImageBGSubtract: single-channel images only')
        return ops.image_pipe_host('bgsubtract', image, out_dtype='float64')


class ImageWeightMap(ImagePipe):
    """ ImageWeightMap (pipeline.py:455-479): exponential decay away from the edges of
    binary objects, w = w0*(1-m)*exp(-d^2/(2 sigma^2)) + m + 1 with d the exact
    Euclidean distance to the nearest foreground pixel.  Runs on the GPU
    (``sq_weightmap_edt``); returns (H,W,1) float64 like the reference. """

    def __init__(self, w0=10., sigma=5.):
        ImagePipe.__init__(self)
        self.w0 = w0
        self.sigma = sigma

    def pipe(self, image):
        m = _binary_plane(image, 'ImageWeightMap')
        w = ops.weightmap_edt_host(m, self.w0, self.sigma, out_dtype='float64')
        return w[..., np.newaxis]


class ImageWeightMapUNet(ImagePipe):
    """ North-star U-Net weight map: w = w_c + w0*exp(-(d1+d2)^2/(2 sigma^2)), d1/d2 the
    distances to the nearest and second-nearest distinct instance.  Input is an
    integer instance-label image; a {0,1}/bool mask is first split into its
    4-connected components on the GPU.  Returns (H,W,1) float64. """

    def __init__(self, w0=10., sigma=5., wc=None):
        ImagePipe.__init__(self)
        self.w0 = w0
        self.sigma = sigma
        self.wc = wc

    def __call__(self, image):
        image = np.asarray(image)
        if image.ndim == 3 and image.shape[-1] == 1:
            image = image[..., 0]
        if image.ndim != 2:
            raise NotImplementedError('ImageWeightMapUNet: (H,W) label images only')
        return self.pipe(image)

    def pipe(self, image):
        if image.dtype == np.bool_ or (image.dtype.kind == 'f') or image.max(initial=0) <= 1:
            binary = (image != 0).astype(np.uint8)
            _, labels = ops.label_centroids_host(binary[None], want_labels=True)
            labels = labels[0]
        else:
            labels = image.astype(np.int32)
        w = ops.weightmap_unet_host(labels, self.w0, self.sigma, self.wc, out_dtype='float64')
        return w[..., np.newaxis]


class ImageWeightMap2(ImagePipe):
    """ ImageWeightMap2 (pipeline.py:482-571): the reference's Delaunay "gap
    narrowness" approximation of the same idea.  HOST-SIDE by design: the result
    depends on Qhull's triangulation and on ``find_simplex`` tie-breaking for pixels
    that lie exactly on triangle edges, which no GPU restatement can reproduce
    bit for bit; it calls the same SciPy routines the reference calls.  It is NOT
    part of the accelerated path -- use ImageWeightMap / ImageWeightMapUNet there. """

    def __init__(self, w0=10., sigma=5.):
        ImagePipe.__init__(self)
        self.w0 = w0
        self.sigma = sigma

    def pipe(self, image):
        from scipy.ndimage import binary_dilation, binary_erosion, gaussian_filter
        from scipy.spatial import Delaunay
        s = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
        b = np.squeeze(image.astype('bool'))
        b_erode_outline = np.logical_xor(binary_erosion(b, iterations=1, structure=s), b)
        b_dilate = binary_dilation(b, iterations=3, structure=s)
        b_dilate_outline = np.logical_xor(binary_erosion(b_dilate, iterations=1, structure=s), b_dilate)
        b_erode = np.logical_xor(b_erode_outline, b_dilate_outline)
        x, y = np.where(b_erode)
        tri = Delaunay(np.column_stack((x, y)))
        self.tri = tri
        free_x, free_y = np.where(np.logical_not(b))
        simplices = tri.find_simplex(np.column_stack((free_x, free_y)))
        pts = tri.points[tri.simplices]
        longest = np.sqrt(((pts - np.roll(pts, -1, axis=1)) ** 2).sum(-1)).max(-1)
        weight_map = np.zeros(image.shape)
        weight_map[free_x, free_y, ...] = np.where(simplices >= 0, longest[np.maximum(simplices, 0)],
                                                   1024.).reshape((-1, 1))
        mask = b[..., np.newaxis].astype('float32')
        weight_map = gaussian_filter(weight_map, 1.)
        weight_map = self.w0 * (1. - mask) * np.exp(-(weight_map * weight_map) /
                                                    (2. * self.sigma ** 2 + 1e-99))
        return weight_map + 1. + mask
