"""Powder characterisation -- drop-in for ``ampis.applications.powder`` (reference
ampis/applications/powder.py) with the satellite/particle overlap assignment and the mask
areas computed on the GPU.  Plotting (matplotlib) is outside the accelerated path: ``psd``
computes and returns the distribution, and draws only if an axis is passed.
"""
import copy

import numpy as np

from .. import analyze, engine
from ..containers import Instances
from ..structures import InstanceSet, mask_areas, masks_to_rle


def _rle_satellite_match(particles, satellites, match_thresh=0.5):
    """Match every satellite to the particle it overlaps most (reference powder.py:28-112).

    score(s, p) = |s AND p| / |s|; satellite s goes to argmax_p (first on ties) if that score is
    strictly above *match_thresh*.  A satellite matches at most one particle, a particle may own
    many satellites.  The reference evaluates all S x N pairs with RLE.merge + RLE.area; the
    fused row kernel visits only pairs whose boxes overlap (the others intersect in 0 pixels,
    which cannot win an arg-max against any positive score and ties resolve to index 0 either way).
    Raises IndexError when no satellite matches, like the reference (powder.py:101)."""
    particles = masks_to_rle(particles)
    satellites = masks_to_rle(satellites)
    S, Np = len(satellites), len(particles)
    if S and Np == 0:
        raise ValueError('attempt to get argmax of an empty sequence')
    if S:
        try:        # the marshaller checks every mask's size on its way (a Python loop over them cost half the call)
            best, score, _, _ = analyze._image_rows(satellites, particles, engine.MODE_SAT)   # NaN score for zero-area
        except ValueError as e:                                                               # satellites, as numpy's 0/0
            if 'different image sizes' in str(e):
                raise ValueError('particle and satellite masks must share one image size') from e
            raise
    else:
        best, score = np.zeros(0, np.int64), np.zeros(0)
    with np.errstate(invalid='ignore'):
        matched = score > match_thresh
    particles_matched_bool = np.zeros(Np, dtype=np.bool_)
    particles_matched_bool[best[matched]] = True
    satellite_matches = np.stack([np.nonzero(matched)[0], best[matched]], axis=1).astype(np.int64) \
        if matched.any() else np.asarray([], np.int64)
    satellites_unmatched = np.nonzero(~matched)[0].astype(np.int64)
    particles_unmatched = np.nonzero(~particles_matched_bool)[0].astype(np.int64)
    intersection_scores = score[matched]

    match_pairs = {x: [] for x in np.unique(satellite_matches[:, 1])}   # IndexError if nothing matched
    for match in satellite_matches:
        match_pairs[match[1]].append(match[0])

    return {'satellite_matches': satellite_matches,
            'satellites_unmatched': satellites_unmatched,
            'particles_unmatched': particles_unmatched,
            'intersection_scores': intersection_scores,
            'match_pairs': match_pairs}


#: legacy name
fast_satellite_match = _rle_satellite_match


class PowderSatelliteImage(object):
    """Particle and satellite instances of one image (reference powder.py:115-285)."""

    def __init__(self, particles=None, satellites=None, matches=None):
        self.particles = particles
        self.satellites = satellites
        self.matches = matches

    def compute_matches(self, thresh=0.5):
        self.matches = _rle_satellite_match(self.particles.instances, self.satellites.instances, thresh)

    def compute_satellite_metrics(self):
        """Counts and particle mask areas for size filtering (reference powder.py:221-273)."""
        assert None not in (self.particles, self.satellites, self.matches)
        n_satellites = len(self.satellites.instances)
        matched_particle_idx = list(self.matches['match_pairs'])
        n_particles_matched = len(matched_particle_idx)
        n_particles_all = len(self.particles.instances)
        particle_masks_all = masks_to_rle(self.particles.instances.masks.rle)
        mask_areas_all = mask_areas(particle_masks_all)
        mask_areas_matched = mask_areas_all[matched_particle_idx]
        return {'n_satellites': n_satellites,
                'n_particles_matched': n_particles_matched,
                'n_particles_all': n_particles_all,
                'mask_areas_matched': mask_areas_matched,
                'mask_areas_all': mask_areas_all}

    def copy(self):
        return copy.deepcopy(self)


def _pixel_scale(isets):
    """Length per pixel of every image from its horizontal field width (HFW / image width), and the
    unit string, for psd(distance='length') without an explicit c (reference powder.py:370-388)."""
    if type(isets[0]) != InstanceSet:
        raise ValueError('Cannot infer c from particles (must be list of InstanceSet or PowderSatelliteImage '
                         'objects')
    if isets[0].HFW is None:
        raise ValueError('Cannot infer c because HFW is not defined')
    widths_um = [s.HFW for s in isets]
    assert None not in widths_um, 'all HFW values must be specified if c is not defined'
    units = isets[0].HFW_units
    assert all(s.HFW_units == units for s in isets), 'all HFW values should have same units'
    widths_px = [int(s.instances.image_size[1]) for s in isets]
    return [um / px for um, px in zip(widths_um, widths_px)], units


_PSD_X = {'d_eq': ('Equivalent diameter', ', {}'), 'area': ('Mask area', '- ${}^2$')}
_PSD_Y = {'cvf': 'cumulative volume fraction', 'counts': 'counts (cumulative)'}


def psd(particles, xvals='d_eq', yvals='cvf', c=None, distance='length', ax=None, plot=True, return_results=False,
        _areas_of=None):
    """Cumulative particle size distribution (reference powder.py:288-461).

    Areas come from the GPU measurement pass; they are scaled to length units by *c* (a number, one
    number per image, or ``(c, unit)``; inferred from HFW when omitted), or left in pixels with
    ``distance='pixels'``.  Distinct areas are counted exactly (``np.unique``), mapped to
    ``d_eq = 2 sqrt(A / pi)`` when ``xvals='d_eq'``, weighted by the sphere volume
    ``4/3 pi^(-1/2) A^(3/2)`` when ``yvals='cvf'``, accumulated and normalised to end at 1.  The same
    arguments raise the same ``ValueError``s as the reference.  Nothing is drawn unless an axis is
    passed (matplotlib is not a dependency); ``return_results=True`` returns the curve and labels.
    ``_areas_of`` (internal): replaces the per-image area measurement -- ``distributed.psd_sharded`` passes the
    sharded, all-gathered form, everything after it is this very code on every rank."""
    if _areas_of is None:
        _areas_of = lambda items: [mask_areas(item) for item in items]
    xkey, ykey, how = xvals.lower(), yvals.lower(), distance.lower()
    units = ''
    if type(c) == tuple:
        c, units = c
    if type(particles) in (InstanceSet, PowderSatelliteImage):
        particles = [particles]
    if type(particles[0]) == PowderSatelliteImage:
        particles = [item.particles for item in particles]
    per_image = _areas_of(particles)                              # the reference's list branch is dead (quirk B.6)

    if how == 'length':
        if c is None:
            c, units = _pixel_scale(particles)
        if type(c) in [list, np.ndarray]:
            assert len(c) == len(per_image), 'if c (or c[0] if passed as tuple) is a list or array ' \
                                             'it must have the same length as particles.'
            per_image = [a * k ** 2 for a, k in zip(per_image, c)]
        elif type(c) in [int, float]:                                # numpy scalars are rejected, as in the reference
            per_image = [a * c ** 2 for a in per_image]
        else:
            raise ValueError('c (or c[0] if passed as tuple) must be a list, array, int, or float')
    elif how == 'pixels':
        units = 'px'
        per_image = _areas_of(particles)                             # recomputed, as in the reference (powder.py:410)
    else:
        raise ValueError('distance must be "length" or "pixels"')
    areas = np.concatenate(per_image, axis=0) if type(per_image[0]) in (list, np.ndarray) else per_image

    x, weight = np.unique(areas, return_counts=True)
    if xkey not in _PSD_X:
        raise ValueError('xvals must be "d_eq" or "area"')
    if xkey == 'd_eq':
        x = 2 * np.sqrt(x / np.pi)
    name, unit_fmt = _PSD_X[xkey]
    xlabel = name + (unit_fmt.format(units) if units else '')
    if ykey not in _PSD_Y:
        raise ValueError('yvals must be "cvf" or "counts"')
    if ykey == 'cvf':
        weight = 4 / 3 * np.pi ** (-1 / 2) * x ** (3 / 2) * weight
    ylabel = _PSD_Y[ykey]
    y = weight.cumsum()
    y = y / y[-1]

    if ax is not None:
        ax.grid(axis='both', which='both', color=(0.85, 0.85, 0.85), linewidth=1, linestyle='--')
        ax.plot(x, y, '-.k')
        ax.set_xlabel(xlabel)
        ax.set_ylabel(ylabel)
    if return_results:
        return {'x': x, 'y': y, 'x_label': xlabel, 'y_label': ylabel}


_SUMMARY_ROWS = (('n_images', 'number of images'),
                 ('n_particles', 'number of particles'),
                 ('n_satellites', 'number of matched satellites'),
                 ('n_satellites_unmatched', 'number of unmatched satellites'),
                 ('n_satellited_particels', 'number of satellited particles'),            # key spelled as in the reference
                 ('sat_frac', 'fraction of satellited particles'),
                 ('mspp', 'median number of satellites per\nsatellited particle             '))


def _satellite_summary(per_particle, n_images, n_particles_unmatched, n_satellites_unmatched, n_particle_instances,
                       n_satellite_instances):
    """The dataset-level numbers of powder.py:525-562 from the satellites-per-satellited-particle list and four
    sums (shared by satellite_measurements and distributed.satellite_measurements_sharded, which gathers exactly
    these quantities from all ranks)."""
    per_particle = np.asarray(per_particle)
    out = {'n_images': n_images}
    satellited = len(per_particle)
    out['n_particles'] = satellited + n_particles_unmatched
    out['n_satellites'] = sum(per_particle)
    out['n_satellites_unmatched'] = n_satellites_unmatched
    out['n_satellited_particels'] = satellited
    out['sat_frac'] = satellited / out['n_particles']
    out['mspp'] = np.median(per_particle)
    values, freq = np.unique(per_particle, return_counts=True)
    # consistency with the instance lists themselves (the reference asserts the same three sums)
    assert freq.sum() == satellited
    assert out['n_particles'] == n_particle_instances
    assert out['n_satellites'] + out['n_satellites_unmatched'] == n_satellite_instances
    out['unique_satellites_per_particle'] = values
    out['counts_satellites_per_particle'] = freq.cumsum() / freq.sum()
    return out


def _print_summary(out):
    for key, label in _SUMMARY_ROWS:
        print('{:35}\t{}'.format(label, out[key]))


def satellite_measurements(psi, print_summary=True, output_dict=False):
    """Satellite content of a set of images (reference powder.py:463-569): totals, the fraction of
    particles that carry satellites, the median number of satellites per satellited particle and
    the cumulative distribution of that number.  Images without matches get them computed first.
    The same sums are what ``distributed.satellites_sharded`` all-reduces over GPUs."""
    images = [psi] if type(psi) == PowderSatelliteImage else psi
    assert all(type(im) == PowderSatelliteImage for im in images), 'psi must be list of PowderSatelliteImage objects!'
    if any(im.matches is None for im in images):
        for im in images:
            im.compute_matches()
    found = [im.matches for im in images]

    per_particle = np.asarray([len(sats) for m in found for sats in m['match_pairs'].values()])
    out = _satellite_summary(per_particle, len(images),
                             sum(len(m['particles_unmatched']) for m in found),
                             sum(len(m['satellites_unmatched']) for m in found),
                             sum(len(im.particles.instances) for im in images),
                             sum(len(im.satellites.instances) for im in images))

    if print_summary:
        _print_summary(out)
    if output_dict:
        return out
