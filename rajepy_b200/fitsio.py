"""
Minimal FITS primary-HDU writer/reader for the model's data products.

The reference writes its images with astropy (`JetModel.save_fits`, classes.py:1543-1652);
astropy is not a dependency here, so this module emits the same primary HDU itself:
float64 big-endian data in the (.., Dec, RA) axis order the caller prepared with
`reorder_axes`, and the same header keywords, values and comments in the same order
(AUTHOR .. CDELT2, the optional FREQ axis, BUNIT and the model table as HISTORY cards).
`read_fits` is a small reader used by the tests (and handy for quick looks).
"""
import numpy as np
import scipy.constants as con

BLOCK = 2880
CARD = 80


def parse_sexagesimal(ra, dec):
    """('HH:MM:SS.S', '+DD:MM:SS.S') -> (ra_deg, dec_deg), what
    SkyCoord(ra, dec, unit=(hourangle, deg), frame='fk5') gives (classes.py:1575-1577)."""
    def split(s):
        s = s.strip()
        sign = -1.0 if s.startswith('-') else 1.0
        parts = [float(p) for p in s.lstrip('+-').replace(' ', ':').split(':')]
        parts += [0.0] * (3 - len(parts))
        return sign, parts
    sr, (h, m, s) = split(ra)
    sd, (d, dm, ds) = split(dec)
    return sr * 15.0 * (h + m / 60.0 + s / 3600.0), sd * (d + dm / 60.0 + ds / 3600.0)


def _fmt_value(v):
    if isinstance(v, bool):
        return ('T' if v else 'F').rjust(20)
    if isinstance(v, (int, np.integer)):
        return str(int(v)).rjust(20)
    if isinstance(v, (float, np.floating)):
        s = repr(float(v)).upper()
        if 'E' in s:
            mant, exp = s.split('E')
            if '.' not in mant:
                mant += '.0'
            s = mant + 'E' + exp
        elif '.' not in s and 'N' not in s:
            s += '.0'
        return s.rjust(20)
    s = "'" + str(v).replace("'", "''").ljust(8) + "'"
    return s.ljust(20)


def _card(key, value=None, comment=None):
    if value is None:
        txt = key.ljust(8) + (comment or '')
        return txt[:CARD].ljust(CARD)
    txt = key.ljust(8) + '= ' + _fmt_value(value)
    if comment:
        txt += ' / ' + comment
    return txt[:CARD].ljust(CARD)


def _history_cards(text):
    return [('HISTORY ' + text[i:i + 72]).ljust(CARD) for i in range(0, max(len(text), 1), 72)]


def build_header(model, data, image_type, freq=None):
    """Header cards of classes.py:1588-1648, in order."""
    p = model.params
    ndims = data.ndim
    if ndims not in (2, 3):
        raise ValueError(f"Unexpected number of data dimensions ({ndims})")
    ra_deg, dec_deg = parse_sexagesimal(p['target']['ra'], p['target']['dec'])
    csize_deg = float(np.degrees(np.arctan(model.csize * con.au /
                                           (p['target']['dist'] * con.parsec))))
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'),
             _card('BITPIX', -64, 'array data type'),
             _card('NAXIS', ndims, 'number of array dimensions')]
    for i, n in enumerate(reversed(data.shape)):
        cards.append(_card(f'NAXIS{i + 1}', int(n)))
    cards += [
        _card('EXTEND', True),
        _card('AUTHOR', 'S.J.D.Purser'),
        _card('OBJECT', p['target']['name']),
        _card('CTYPE1', 'RA---TAN', 'x-coord type is RA Tan Gnomonic projection'),
        _card('CTYPE2', 'DEC--TAN', 'y-coord type is DEC Tan Gnomonic projection'),
        _card('EQUINOX', 2000., 'Equinox of coordinates'),
        _card('CRPIX1', model.nx / 2 + 0.5, 'Reference pixel in RA'),
        _card('CRPIX2', model.nz / 2 + 0.5, 'Reference pixel in DEC'),
        _card('CRVAL1', ra_deg, 'Reference pixel value in RA (deg)'),
        _card('CRVAL2', dec_deg, 'Reference pixel value in DEC (deg)'),
        _card('CDELT1', -csize_deg, 'Pixel increment in RA (deg)'),
        _card('CDELT2', csize_deg, 'Pixel size in DEC (deg)'),
    ]
    if image_type in ('flux', 'tau', 'intensity'):
        fr = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        if ndims == 3:
            nchan = len(fr)
            chan_width = float(fr[1] - fr[0]) if nchan != 1 else 1.
            cards += [_card('CTYPE3', 'FREQ', 'Spectral axis (frequency)'),
                      _card('CRPIX3', nchan / 2. + 0.5, 'Reference frequency (channel number)'),
                      _card('CRVAL3', float(fr[len(fr) // 2 - 1] + chan_width / 2),
                            'Reference frequency (Hz)'),
                      _card('CDELT3', chan_width, 'Frequency increment (Hz)')]
        else:
            cards += [_card('CDELT3', 1., 'Frequency increment (Hz)'),
                      _card('CRPIX3', 0.5, 'Reference frequency (channel number)'),
                      _card('CRVAL3', float(fr[0]), 'Reference frequency (Hz)')]
    bunit = {'flux': 'Jy pixel^-1', 'intensity': 'W m^-2 Hz^-1 sr^-1', 'em': 'pc cm^-6',
             'tau': 'dimensionless'}[image_type]
    cards.append(_card('BUNIT', bunit))
    s_hist = str(model).split('\n')
    cards += _history_cards((' ' * (72 - len(s_hist[0]))).join(s_hist))
    cards.append('END'.ljust(CARD))
    return cards


def write_model_fits(model, data, filename, image_type, freq=None):
    data = np.asarray(data, dtype=np.float64)
    cards = build_header(model, data, image_type, freq)
    hdr = ''.join(cards)
    hdr += ' ' * (-len(hdr) % BLOCK)
    flat = np.ascontiguousarray(data).reshape(-1)
    with open(filename, 'wb') as f:
        f.write(hdr.encode('ascii'))
        step = 1 << 23                    # 64 MB of big-endian doubles at a time
        for i in range(0, flat.size, step):
            f.write(flat[i:i + step].astype('>f8').tobytes())
        f.write(b'\0' * (-(flat.size * 8) % BLOCK))


def read_fits_data(filename):
    """Data array of a primary HDU (what `fits.open(f)[0].data` gives, classes.py:2431)."""
    return read_fits(filename)[1]


def read_fits(filename):
    """-> (header dict incl. 'HISTORY' list, data ndarray) of a primary HDU."""
    with open(filename, 'rb') as f:
        blob = f.read()
    hdr, pos, done = {'HISTORY': []}, 0, False
    while not done:
        block = blob[pos:pos + BLOCK].decode('ascii')
        pos += BLOCK
        for i in range(0, BLOCK, CARD):
            card = block[i:i + CARD]
            key = card[:8].strip()
            if key == 'END':
                done = True
                break
            if key == 'HISTORY':
                hdr['HISTORY'].append(card[8:])
            elif card[8:10] == '= ':
                val = card[10:].split(' / ')[0].strip()
                if val.startswith("'"):
                    hdr[key] = val.strip("'").rstrip()
                elif val in ('T', 'F'):
                    hdr[key] = val == 'T'
                else:
                    hdr[key] = float(val) if ('.' in val or 'E' in val) else int(val)
    shape = tuple(hdr[f'NAXIS{i}'] for i in range(hdr['NAXIS'], 0, -1))
    n = int(np.prod(shape))
    data = np.frombuffer(blob, dtype='>f8', count=n, offset=pos).reshape(shape)
    return hdr, data.astype(np.float64)
