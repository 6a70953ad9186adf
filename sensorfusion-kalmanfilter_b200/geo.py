"""Latitude / longitude -> UTM for the CSV front end (`gps_to_modified_utm`, kf_workers.py:304-332,
hw5_2.py:28-52).

The reference calls the third-party package `utm` (`utm.from_latlon`, pinned nowhere; 0.8.1 at the
time of writing), which is not vendored in the reference and not installed here.  This module
restates that package's published algorithm -- the classic truncated series on the WGS-84 ellipsoid
(Snyder, "Map Projections: A Working Manual", USGS PP 1395, eqs. 8-9 / 8-10 with k0 = 0.9996) -- so
that the front end runs without it.  **Parity unpinned**: no output of `utm` is available in this
repository to compare against; the tests check the zone of the reference's data set (19 T,
KF_SensorFusion.ipynb:1331), the projection's defining properties and the round trip.  It is
upstream of the hot path: the filters only ever see eastings / northings relative to the first fix.
"""
from __future__ import annotations

import math

K0 = 0.9996
E = 0.00669438          # first eccentricity squared (WGS-84, as the package rounds it)
E2 = E * E
E3 = E2 * E
E_P2 = E / (1.0 - E)
SQRT_E = math.sqrt(1.0 - E)
_E = (1.0 - SQRT_E) / (1.0 + SQRT_E)
_E2 = _E * _E
_E3 = _E2 * _E
_E4 = _E3 * _E
_E5 = _E4 * _E
M1 = 1.0 - E / 4.0 - 3.0 * E2 / 64.0 - 5.0 * E3 / 256.0
M2 = 3.0 * E / 8.0 + 3.0 * E2 / 32.0 + 45.0 * E3 / 1024.0
M3 = 15.0 * E2 / 256.0 + 45.0 * E3 / 1024.0
M4 = 35.0 * E3 / 3072.0
P2 = 3.0 / 2.0 * _E - 27.0 / 32.0 * _E3 + 269.0 / 512.0 * _E5
P3 = 21.0 / 16.0 * _E2 - 55.0 / 32.0 * _E4
P4 = 151.0 / 96.0 * _E3 - 417.0 / 128.0 * _E5
P5 = 1097.0 / 512.0 * _E4
R = 6378137.0
ZONE_LETTERS = "CDEFGHJKLMNPQRSTUVWXX"


def latitude_to_zone_letter(latitude):
    if -80.0 <= latitude <= 84.0:
        return ZONE_LETTERS[int(latitude + 80.0) >> 3]
    return None


def latlon_to_zone_number(latitude, longitude):
    if 56.0 <= latitude < 64.0 and 3.0 <= longitude < 12.0:  # Norway
        return 32
    if 72.0 <= latitude <= 84.0 and longitude >= 0.0:        # Svalbard
        if longitude < 9.0:
            return 31
        if longitude < 21.0:
            return 33
        if longitude < 33.0:
            return 35
        if longitude < 42.0:
            return 37
    return int((longitude + 180.0) / 6.0) % 60 + 1


def zone_number_to_central_longitude(zone_number):
    return (zone_number - 1) * 6 - 180 + 3


def _mod_angle(value):
    """Wrap to (-pi, pi]."""
    return (value + math.pi) % (2.0 * math.pi) - math.pi


def from_latlon(latitude, longitude, force_zone_number=None):
    """(easting, northing, zone_number, zone_letter) of a WGS-84 latitude / longitude in degrees."""
    if not -80.0 <= latitude <= 84.0:
        raise ValueError("latitude out of range (must be between 80 deg S and 84 deg N)")
    if not -180.0 <= longitude <= 180.0:
        raise ValueError("longitude out of range (must be between 180 deg W and 180 deg E)")
    lat_rad = math.radians(latitude)
    lat_sin, lat_cos = math.sin(lat_rad), math.cos(lat_rad)
    lat_tan = lat_sin / lat_cos
    lat_tan2 = lat_tan * lat_tan
    lat_tan4 = lat_tan2 * lat_tan2
    zone_number = latlon_to_zone_number(latitude, longitude) if force_zone_number is None else force_zone_number
    zone_letter = latitude_to_zone_letter(latitude)
    lon_rad = math.radians(longitude)
    central_lon_rad = math.radians(zone_number_to_central_longitude(zone_number))
    n = R / math.sqrt(1.0 - E * lat_sin ** 2)
    c = E_P2 * lat_cos ** 2
    a = lat_cos * _mod_angle(lon_rad - central_lon_rad)
    a2 = a * a
    a3 = a2 * a
    a4 = a3 * a
    a5 = a4 * a
    a6 = a5 * a
    m = R * (M1 * lat_rad - M2 * math.sin(2.0 * lat_rad) + M3 * math.sin(4.0 * lat_rad) - M4 * math.sin(6.0 * lat_rad))
    easting = K0 * n * (a + a3 / 6.0 * (1.0 - lat_tan2 + c) +
                        a5 / 120.0 * (5.0 - 18.0 * lat_tan2 + lat_tan4 + 72.0 * c - 58.0 * E_P2)) + 500000.0
    northing = K0 * (m + n * lat_tan * (a2 / 2.0 + a4 / 24.0 * (5.0 - lat_tan2 + 9.0 * c + 4.0 * c ** 2) +
                                        a6 / 720.0 * (61.0 - 58.0 * lat_tan2 + lat_tan4 + 600.0 * c - 330.0 * E_P2)))
    if latitude < 0.0:
        northing += 10000000.0
    return easting, northing, zone_number, zone_letter


def to_latlon(easting, northing, zone_number, zone_letter=None, northern=None):
    """Inverse series (used by the tests for the round trip)."""
    if northern is None:
        northern = zone_letter is None or zone_letter.upper() >= "N"
    x = easting - 500000.0
    y = northing if northern else northing - 10000000.0
    m = y / K0
    mu = m / (R * M1)
    p_rad = (mu + P2 * math.sin(2.0 * mu) + P3 * math.sin(4.0 * mu) + P4 * math.sin(6.0 * mu) + P5 * math.sin(8.0 * mu))
    p_sin, p_cos = math.sin(p_rad), math.cos(p_rad)
    p_tan = p_sin / p_cos
    p_tan2 = p_tan * p_tan
    p_tan4 = p_tan2 * p_tan2
    ep_sin = 1.0 - E * p_sin * p_sin
    ep_sin_sqrt = math.sqrt(ep_sin)
    n = R / ep_sin_sqrt
    r = (1.0 - E) / ep_sin
    c = E_P2 * p_cos ** 2
    c2 = c * c
    d = x / (n * K0)
    d2 = d * d
    d3 = d2 * d
    d4 = d3 * d
    d5 = d4 * d
    d6 = d5 * d
    latitude = p_rad - (p_tan / r) * (d2 / 2.0 - d4 / 24.0 * (5.0 + 3.0 * p_tan2 + 10.0 * c - 4.0 * c2 - 9.0 * E_P2) +
                                      d6 / 720.0 * (61.0 + 90.0 * p_tan2 + 298.0 * c + 45.0 * p_tan4 - 252.0 * E_P2 - 3.0 * c2))
    longitude = (d - d3 / 6.0 * (1.0 + 2.0 * p_tan2 + c) +
                 d5 / 120.0 * (5.0 - 2.0 * c + 28.0 * p_tan2 - 3.0 * c2 + 8.0 * E_P2 + 24.0 * p_tan4)) / p_cos
    longitude = _mod_angle(longitude + math.radians(zone_number_to_central_longitude(zone_number)))
    return math.degrees(latitude), math.degrees(longitude)
