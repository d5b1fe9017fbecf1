fs = 0.09822694788464063  # ASE time unit: 1 fs in sqrt(u * A^2 / eV)
GPa = 0.006241509125883258  # eV/A^3 per GPa
kB = 8.617330337217213e-05
bar = 1.0e-4 * GPa
