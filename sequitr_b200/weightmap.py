"""Drop-in for the reference ``sequitr/weightmap.py``: ``ImageLabels`` (:31-73),
``create_weightmaps`` (:171-205) and the CLI (:210-233).

``create_weightmaps(method=...)`` selects the weight pipe:
  'delaunay' (default) the reference's own choice, ``pipeline.ImageWeightMap2``
             (weightmap.py:181; host, SciPy/Qhull) -- the same call writes the same files;
  'edt'      opt-in: GPU ``pipeline.ImageWeightMap`` (exact EDT, single distance);
  'unet'     opt-in: GPU north-star map w_c + w0*exp(-(d1+d2)^2/2 sigma^2); the bool
             mask the reference passes (weightmap.py:203) is split into instances
             by the GPU connected-component kernel first.
The two GPU methods write numerically DIFFERENT maps than the reference's Delaunay
approximation, which is why they are never the default.
"""
import os
import re

import numpy as np

from .pipeline import ImagePipeline, ImageWeightMap, ImageWeightMap2, ImageWeightMapUNet  # noqa: F401


def _imread(filename):
    try:
        import tifffile
        return tifffile.imread(filename)
    except ImportError:
        import cv2
        ok, pages = cv2.imreadmulti(filename, flags=cv2.IMREAD_UNCHANGED)
        if not ok or not pages:
            raise IOError('cannot read %s' % filename)
        return pages[0] if len(pages) == 1 else np.stack(pages)


def _imsave(filename, data):
    try:
        import tifffile
        tifffile.imsave(filename, data)
    except ImportError:
        import cv2
        if not cv2.imwrite(filename, data):
            raise IOError('cannot write %s' % filename)


def check_and_makedir(folder_name):
    """ utils.check_and_makedir """
    if not os.path.exists(folder_name):
        os.makedirs(folder_name)
        return False
    return True


class ImageLabels(object):
    """ ImageLabels (weightmap.py:31-73): TIFF -> uint8 label image. """

    def __init__(self, filename, thresh_fn=lambda x: x > 0):
        self._raw_data = filename if isinstance(filename, np.ndarray) else _imread(filename)
        assert(self._raw_data.ndim > 1 and self._raw_data.ndim < 4)
        if self._raw_data.ndim == 3:
            l_data = np.zeros(self._raw_data.shape[1:], dtype='uint8')
            for l in range(self._raw_data.shape[0]):
                l_data[thresh_fn(self._raw_data[l, ...])] = l + 1
            raw_labels = range(self._raw_data.shape[0] + 1)
        else:
            l_data = thresh_fn(self._raw_data).astype('uint8')
            raw_labels = [0, 1]
        self._outputs = len(raw_labels)
        if self.outputs > 5:
            raise ValueError('More that five output classes!')
        self._labels = l_data

    def labels(self):
        """ return the labels """
        return self._labels

    @property
    def outputs(self):
        return self._outputs


_METHODS = {'unet': ImageWeightMapUNet, 'edt': ImageWeightMap, 'delaunay': ImageWeightMap2}


def create_weightmaps(path, folders, w0=10., sigma=3., thresh_fn=lambda x: x > 0,
                      name_weights_folder=True, method='delaunay'):
    """ Generate weightmaps for the images using the binary masks (weightmap.py:171-205) """
    if method not in _METHODS:
        raise ValueError('method must be one of %s' % sorted(_METHODS))
    w_pipe = _METHODS[method](w0=w0, sigma=sigma)
    written = []
    for d in folders:
        r_dir = os.path.join(path, d)
        f_labels = [l for l in os.listdir(os.path.join(r_dir, 'label/')) if l.endswith('.tif')]
        w_dir_base = 'weights'
        if name_weights_folder:
            w_dir_base += '_w0-{0:2.2f}_sigma-{1:2.2f}'.format(w0, sigma)
        w_dir = os.path.join(r_dir, w_dir_base)
        check_and_makedir(w_dir)
        for f in sorted(f_labels):
            w_label = re.match('([a-zA-Z0-9()]+)_([a-zA-Z0-9()]+_)*', f).group(0)
            w_label += 'weights.tif'
            label_filename = os.path.join(r_dir, 'label/', f)
            im_label = ImageLabels(label_filename, thresh_fn).labels()
            im_weights = np.squeeze(w_pipe(im_label.astype('bool')))
            _imsave(os.path.join(w_dir, w_label), im_weights.astype('float32'))
            written.append(os.path.join(w_dir, w_label))
    return written


def main(argv=None):
    import argparse
    p = argparse.ArgumentParser(description='Sequitr: weightmap calculation')
    p.add_argument('-p', '--workdir', required=True, help='Path to the image data')
    p.add_argument('-f', '--folders', nargs='+', required=True,
                   help='Specify the sub-folders of image data')
    p.add_argument('--w0', type=float, default=30., help='Specify the amplitude')
    p.add_argument('--sigma', type=float, default=3., help='Specify the sigma')
    p.add_argument('--method', default='delaunay', choices=sorted(_METHODS),
                   help="'delaunay' = the reference's ImageWeightMap2 (host); 'edt' / 'unet' = GPU maps (opt-in)")
    args = p.parse_args(argv)
    create_weightmaps(args.workdir, args.folders, w0=args.w0, sigma=args.sigma, method=args.method)


if __name__ == '__main__':
    main()
