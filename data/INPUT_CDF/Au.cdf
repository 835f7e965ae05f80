Gold
1		! how many elements are in this molecule of solid
79	1	! atomic number, contribution of 1st element into compount 
19.32	2030.0	5.53	! density [g/cm^3], speed of sound [m/s], Fermi energy [eV]
6		! number of shells of the first element: Au
1	1	80700.0d0	2	0.392d0	! number of CDF functions, shell-designator, ionization potential, number of electrons, Auger-time
82000	230	50000	! E0, A, Gamma coefficients

1	2	11400.0d0	8	0.245d0	! number of CDF functions, shell-designator, ionization potential, number of electrons, Auger-time
14000	260	12000	! E0, A, Gamma coefficients

1	7	2066.0d0	18	0.487d0	! number of CDF functions, shell-designator, ionization potential, number of electrons, Auger-time
2550	1700	1650	! E0, A, Gamma coefficients

2	15	390.0d0	18	2.88d0	! number of CDF functions, shell-designator, ionization potential, number of electrons, Auger-time
450	750	250
800	900	500	! E0, A, Gamma coefficients

2	26	55.0d0	22	2.88d0	! number of CDF functions, shell-designator, ionization potential, number of electrons, Auger-time
75	1000	65
250	900	250	! E0, A, Gamma coefficients

9	63	0.0d0	11	1.0d23	! number of CDF functions, shell-designator, ionization potential, number of electrons, Auger-time
1.0	5.0e-4	0.9
1.9	-0.14	2.5
2.0	-0.042	0.9
2.8	0.23	0.8
9.0	10.0	9.0
16.2	62.0	9.0
25.7	55.0	5.0
35.0	380.0	14
46.0	360.0	15.0	! E0, A, Gamma coefficients
