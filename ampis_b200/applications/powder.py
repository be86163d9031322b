"""Powder characterisation -- drop-in for ``ampis.applications.powder`` (reference
ampis/applications/powder.py) with the satellite/particle overlap assignment and the mask
areas computed on the GPU.  Plotting (matplotlib) is outside the accelerated path: ``psd``
computes and returns the distribution, and draws only if an axis is passed.
"""
import copy

import numpy as np

from .. import engine
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
        for m in list(particles) + list(satellites):
            if list(m['size']) != list(particles[0]['size']):
                raise ValueError('particle and satellite masks must share one image size')
        table = engine.table_from_rle(list(satellites) + list(particles), layout=engine.MATCH_LAYOUT)
        groups = engine.Groups.interleaved(table.device, [S], [Np])
        res = engine.intersect_rows(table, groups, engine.MODE_SAT)
        best = res.best_col[:S].cpu().numpy().astype(np.int64)
        score = res.best_score[:S].cpu().numpy()          # NaN for zero-area satellites, as numpy's 0/0
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


def psd(particles, xvals='d_eq', yvals='cvf', c=None, distance='length', ax=None, plot=True, return_results=False):
    """Cumulative particle size distribution from segmentation masks (reference
    powder.py:288-461).  Same argument handling and ValueErrors; the numerics (exact-value
    histogram via np.unique, d_eq = 2*sqrt(A/pi), V = 4/3*pi^(-1/2)*A^(3/2), normalised cumsum)
    follow powder.py:417-444 on areas measured on the GPU.  Drawing needs matplotlib and happens
    only on an axis passed by the caller."""
    if type(c) == tuple:
        length_units = c[1]
        c = c[0]
    else:
        length_units = ''
    if type(particles) in (InstanceSet, PowderSatelliteImage):
        particles = [particles]
    if type(particles[0]) == PowderSatelliteImage:
        particles = [x.particles for x in particles]
    areas = [mask_areas(x) for x in particles]      # powder.py:363 is always truthy (quirk B.6)

    if distance.lower() == 'length':
        if c is None:
            if type(particles[0]) == InstanceSet:
                if particles[0].HFW is not None:
                    HFW = [x.HFW for x in particles]
                    assert all([x is not None for x in HFW]), 'all HFW values must be specified if c is not defined'
                    for iset in particles:
                        assert iset.HFW_units == particles[0].HFW_units, 'all HFW values should have same units'
                    length_units = particles[0].HFW_units
                    HFW = np.asarray([x.HFW for x in particles])
                    image_widths = np.asarray([x.instances.image_size[1] for x in particles], np.int64)
                    c = [h / w for h, w in zip(HFW, image_widths)]
                else:
                    raise ValueError('Cannot infer c because HFW is not defined')
            else:
                raise ValueError('Cannot infer c from particles (must be list of InstanceSet or PowderSatelliteImage '
                                 'objects')
        if type(c) in [list, np.ndarray]:
            assert len(c) == len(areas), 'if c (or c[0] if passed as tuple) is a list or array ' \
                                         'it must have the same length as particles.'
            areas = [a_i * c_i ** 2 for a_i, c_i in zip(areas, c)]
        elif type(c) in [int, float]:
            areas = [a_i * c ** 2 for a_i in areas]
        else:
            raise ValueError('c (or c[0] if passed as tuple) must be a list, array, int, or float')
    elif distance.lower() == 'pixels':
        length_units = 'px'
        areas = mask_areas(particles)
    else:
        raise ValueError('distance must be "length" or "pixels"')

    if type(areas[0]) in (list, np.ndarray):
        areas = np.concatenate(areas, axis=0)

    unique, counts = np.unique(areas, return_counts=True)
    if xvals.lower() == 'd_eq':
        unique = 2 * np.sqrt(unique / np.pi)
        xlabel = 'Equivalent diameter{}'.format(', {}'.format(length_units) if length_units else '')
    elif xvals.lower() == 'area':
        xlabel = 'Mask area{}'.format('- ${}^2$'.format(length_units) if length_units else '')
    else:
        raise ValueError('xvals must be "d_eq" or "area"')

    if yvals.lower() == 'cvf':
        volumes = 4 / 3 * np.pi ** (-1 / 2) * unique ** (3 / 2)
        counts = volumes * counts
        ylabel = 'cumulative volume fraction'
    elif yvals.lower() == 'counts':
        ylabel = 'counts (cumulative)'
    else:
        raise ValueError('yvals must be "cvf" or "counts"')

    counts = counts.cumsum()
    counts = counts / counts[-1]
    x, y = unique, counts

    if ax is not None:
        ax.grid(axis='both', which='both', color=(0.85, 0.85, 0.85), linewidth=1, linestyle='--')
        ax.plot(x, y, '-.k')
        ax.set_xlabel(xlabel)
        ax.set_ylabel(ylabel)

    if return_results:
        return {'x': x, 'y': y, 'x_label': xlabel, 'y_label': ylabel}


def satellite_measurements(psi, print_summary=True, output_dict=False):
    """Satellite content of a list of PowderSatelliteImage objects (reference powder.py:463-569)."""
    if type(psi) == PowderSatelliteImage:
        psi = [psi]
    assert all([type(x) == PowderSatelliteImage for x in psi]), 'psi must be list of PowderSatelliteImage objects!'
    matches = [x.matches for x in psi]
    if any([x is None for x in matches]):
        for x in psi:
            x.compute_matches()
        matches = [x.matches for x in psi]

    n_images = len(psi)
    n_particles_matched = sum([len(x['match_pairs'].keys()) for x in matches])
    n_particles = n_particles_matched + sum([len(x['particles_unmatched']) for x in matches])
    spp_list = []
    for m in matches:
        for v in m['match_pairs'].values():
            spp_list.append(len(v))
    spp_list = np.asarray(spp_list)
    n_satellites_matched = sum(spp_list)
    mspp = np.median(spp_list)
    n_satellites_unmatched = sum([len(x['satellites_unmatched']) for x in matches])
    sat_frac = n_particles_matched / n_particles
    unique, counts = np.unique(spp_list, return_counts=True)
    assert counts.sum() == n_particles_matched
    assert n_particles == sum([len(x.particles.instances) for x in psi])
    assert n_satellites_matched + n_satellites_unmatched == sum([len(x.satellites.instances) for x in psi])
    counts = counts.cumsum() / counts.sum()

    keys = ['n_images', 'n_particles', 'n_satellites', 'n_satellites_unmatched', 'n_satellited_particels',
            'sat_frac', 'mspp', 'unique_satellites_per_particle', 'counts_satellites_per_particle']
    labels = ['number of images',
              'number of particles',
              'number of matched satellites',
              'number of unmatched satellites',
              'number of satellited particles',
              'fraction of satellited particles',
              'median number of satellites per\n'
              'satellited particle             ']
    values = [n_images, n_particles, n_satellites_matched, n_satellites_unmatched, n_particles_matched,
              sat_frac, mspp, unique, counts]
    if print_summary:
        for lab, v in zip(labels, values[:-2]):
            print('{:35}\t{}'.format(lab, v))
    if output_dict:
        return dict(zip(keys, values))
