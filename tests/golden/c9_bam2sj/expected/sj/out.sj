###STRAND 0:undefined, 1:+, 2:-
###ANNO 0:novel, 1:annotated
###MOTIF 0:non-canonical, 1:GT/AG, 2:CT/AC, 3:GC/AG, 4:CT/GC, 5:AT/AC, 6:GT/AT
#CHR	START	END	STRAND	ANNO	UNIQ_C	MULTI_C	MOTIF
chrA	150	249	1	1	2	1	1
chrA	282	481	0	1	1	0	0
chrA	460	859	1	1	1	0	3
chrA	540	839	0	1	0	1	0
chrB	150	249	2	1	1	1	2
chrB	300	302	0	1	0	1	0
