!======================================================================!
! h9_ref_driver -- TEST INFRASTRUCTURE, UNCOMPILED IN THIS IMAGE (no Fortran
! compiler exists here; see oracle/fortran/Makefile).
!
! Drives the reference's own, unmodified HYDROLOGY.f90 and GROW.f90 through
! the loop nest of PROGRAM H9 (HYBRID9.f90:120-295) on the flat-file data set
! that tools/write_dataset.py writes and hybrid9_b200/host_cpp/h9_driver
! reads, and dumps the same outputs with the prefix out_ref_.  With it the
! C++ oracle (and through it the GPU) can be pinned against the real
! reference on the day a Fortran compiler is available:
!     make -C oracle/fortran            (compiles /root/reference/SOURCE in place)
!     oracle/_ref/h9_ref_driver <data_dir>
!     python tools/compare_ref.py <data_dir>
!
! What is taken from INIT.f90 instead of calling INIT (which needs MPI, netCDF
! and the /scratch data): the allocations :56-135,301-395, dt :214, dz/zc
! :252-257, sla :154, the initial state :707-811 and time_BOY :844-859.
! smp is zeroed (INIT.f90:109 leaves it undefined) and saved/restored per cell
! so that the cell-outer loop order gives the per-cell semantics DESIGN.md
! section 2 defines; set LEAK = .TRUE. to reproduce the reference's shared smp.
!======================================================================!
PROGRAM H9_REF_DRIVER
USE CONTROL
USE SHARED
IMPLICIT NONE
CHARACTER (LEN = 400) :: dir, line
CHARACTER (LEN = 4) :: tag
CHARACTER (LEN = 8), PARAMETER :: fname (7) = (/ 'tas     ','rlds    ','rsds    ', &
                                   'huss    ','ps      ','pr      ','rhs     ' /)
LOGICAL, PARAMETER :: LEAK = .FALSE.
REAL, ALLOCATABLE :: smp_cell (:,:,:)
REAL :: npp_sum, plant_mass_sum, h2osoi_sum_total, decay
INTEGER :: u, n
!----------------------------------------------------------------------!
CALL GET_COMMAND_ARGUMENT (1, dir)
my_id = 0
!----------------------------------------------------------------------!
! driver.txt, positional like INIT.f90:181-206.
!----------------------------------------------------------------------!
ALLOCATE (zi (0:Nlevgrnd), dz (1:Nlevgrnd), zc (1:Nlevgrnd), zc_o (1:nsoil_layers_max))
OPEN (10, FILE = TRIM (dir) // '/driver.txt', STATUS = 'OLD')
READ (10,*) PATH_output
READ (10,*) NISURF
READ (10,*) PGF
READ (10,*) iDEC_start
READ (10,*) iDEC_end
READ (10,*) INTERACTIVE
READ (10,*) LCLIM
READ (10,*) LCLIM_filename
READ (10,*) LSOIL_filename
READ (10,*) syr
READ (10,*) eyr
READ (10,*) NYR_SPIN_UP
READ (10,*) lon_w
READ (10,*) lat_w
READ (10,*) lon_c_w
READ (10,*) lat_c_w
DO I = 0, Nlevgrnd
  READ (10,*) zi (I)
END DO
CLOSE (10)
INTERACTIVE = .FALSE.
OPEN (10, FILE = TRIM (dir) // '/grid.txt', STATUS = 'OLD')
READ (10,*) lon_c, lat_c
CLOSE (10)
lon_s = 1
lat_s = 1
!----------------------------------------------------------------------!
dt = 86400.0 / FLOAT (NISURF)                       ! INIT.f90:214
DO I = 1, Nlevgrnd                                   ! INIT.f90:252-257
  dz (I) = zi (I) - zi (I-1)
END DO
DO I = 1, Nlevgrnd
  zc (I) = zi (I) - dz (I) / 2.0
END DO
DO I = 1, nsoil_layers_max
  zc_o (I) = zc (I)
END DO
IF (iDEC_end < 12) THEN                              ! INIT.f90:289-293
  NYR = (iDEC_end - iDEC_start + 1) * 10
ELSE
  NYR = (iDEC_end - iDEC_start + 1 - 1) * 10 + 2
END IF
!----------------------------------------------------------------------!
! Allocations of INIT.f90:56-135,301-395 that HYDROLOGY/GROW touch.
!----------------------------------------------------------------------!
ALLOCATE (sla (nGPTs), theta_sum (nsoil_layers_max))
n = nsoil_layers_max
ALLOCATE (qin (n+1), qout (n+1), dsmpdw (n+1), dqodw1 (n+1), dqodw2 (n+1), dhkdw (n))
ALLOCATE (dqidw0 (n+1), dqidw1 (n+1), amx (n+1), bmx (n+1), cmx (n+1), rmx (n+1))
ALLOCATE (dwat2 (n+1), dwat (n), GAM (n+1), smp (n), zq (n+1), theta (n), theta_ma (n))
ALLOCATE (S (n), vol_eq (n+1), eff_porosity (n), hk (n), rnff (n+1))
ALLOCATE (plant_mass (nplants_max,lon_c,lat_c), plant_foliage_mass (nplants_max,lon_c,lat_c))
ALLOCATE (plant_length (nplants_max,lon_c,lat_c), rdepth (nplants_max,lon_c,lat_c))
ALLOCATE (nplants (lon_c,lat_c), LAI (lon_c,lat_c), LAI_litter (lon_c,lat_c))
ALLOCATE (rootr_col (1:Nlevgrnd,lon_c,lat_c), soil_tex (lon_c,lat_c), Fmax (lon_c,lat_c))
ALLOCATE (lon (lon_c), lat (lat_c))
ALLOCATE (h2osoi_liq (n,lon_c,lat_c), h2osoi_liq_ma (n,lon_c,lat_c), zwt (lon_c,lat_c), wa (lon_c,lat_c))
ALLOCATE (theta_s (n,lon_c,lat_c), theta_ma_s (n,lon_c,lat_c), hksat (n,lon_c,lat_c))
ALLOCATE (lambda (n,lon_c,lat_c), bsw (n,lon_c,lat_c), psi_s (n,lon_c,lat_c), theta_m (n,lon_c,lat_c))
ALLOCATE (axy_npp (lon_c,lat_c,NYR), axy_plant_mass (lon_c,lat_c,NYR), axy_rnf (lon_c,lat_c,NYR))
ALLOCATE (axy_evap (lon_c,lat_c,NYR), axy_theta (n,lon_c,lat_c,NYR), axy_theta_total (lon_c,lat_c,NYR))
ALLOCATE (smp_cell (n,lon_c,lat_c))
lon (:) = 0.0 ; lat (:) = 0.0
sla (1) = 23.0E-3                                    ! INIT.f90:154
!----------------------------------------------------------------------!
! Soil fields as INIT.f90:470-680 leaves them (flat files, stream access).
!----------------------------------------------------------------------!
CALL READ_I (TRIM (dir) // '/soil_tex.i32', soil_tex, lon_c * lat_c)
CALL READ_R (TRIM (dir) // '/theta_s.f32', theta_s, n * lon_c * lat_c)
CALL READ_R (TRIM (dir) // '/hksat.f32', hksat, n * lon_c * lat_c)
CALL READ_R (TRIM (dir) // '/bsw.f32', bsw, n * lon_c * lat_c)
CALL READ_R (TRIM (dir) // '/psi_s.f32', psi_s, n * lon_c * lat_c)
CALL READ_R (TRIM (dir) // '/fmax.f32', Fmax, lon_c * lat_c)
lambda (:,:,:) = 1.0 / bsw (:,:,:)
theta_ma_s (:,:,:) = 0.1
!----------------------------------------------------------------------!
! Initial state, INIT.f90:707-811.
!----------------------------------------------------------------------!
theta_m (:,:,:) = zero ; h2osoi_liq (:,:,:) = zero ; h2osoi_liq_ma (:,:,:) = zero
plant_mass (:,:,:) = zero ; plant_foliage_mass (:,:,:) = zero ; plant_length (:,:,:) = zero
rdepth (:,:,:) = zero ; zwt (:,:) = zero ; wa (:,:) = zero ; LAI (:,:) = zero
LAI_litter (:,:) = zero ; nplants (:,:) = 0 ; rootr_col (:,:,:) = zero ; smp_cell (:,:,:) = zero
nlayers = nsoil_layers_max
DO y = 1, lat_c
  DO x = 1, lon_c
    IF ((soil_tex (x,y) > 0) .AND. (soil_tex (x,y) /= 13) .AND. &
      SUM (theta_s (:,x,y)) > trunc) THEN
      DO I = 1, nsoil_layers_max
        h2osoi_liq (I,x,y) = 0.4 * theta_s (I,x,y) * dz (I) * rhow / 1000.0
        h2osoi_liq_ma (I,x,y) = 0.4 * theta_ma_s (I,x,y) * dz (I) * rhow / 1000.0
      END DO
      zwt (x,y) = (zi (nlayers) + 5000.0) / 1000.0
      wa (x,y) = 4000.0
      LAI_litter (x,y) = 0.001
      nplants (x,y) = 1
      LAI (x,y) = zero
      rootr_col (:,x,y) = zero
      DO K = 1, nplants (x,y)
        iGPT = 1
        plant_mass (K,x,y) = 1.0
        plant_foliage_mass (K,x,y) = 0.0435
        plant_length (K,x,y) = (400.0 * plant_mass (K,x,y) / 3.142E-3) ** (one / 3.0)
        LAI (x,y) = LAI (x,y) + plant_foliage_mass (K,x,y) * sla (iGPT) / plot_area
        rdepth (K,x,y) = 0.3 * plant_length (K,x,y)
        decay = EXP (LOG (0.1) / (rdepth (K,x,y) / 10.0))
        DO I = 1, nlayers
          rootr_col (I,x,y) = rootr_col (I,x,y) + (1.0 - decay ** (zi (I) / 10.0)) - &
                              (1.0 - decay ** (zi (I-1) / 10.0))
        END DO
      END DO
    END IF
  END DO
END DO
!----------------------------------------------------------------------!
! time_BOY, INIT.f90:844-859.
!----------------------------------------------------------------------!
time_BOY (1) = 1
DO jyear = 1861, 2300
  IF (MOD (jyear-1,4) .NE. 0) THEN
    time_BOY (jyear-1859) = time_BOY (jyear-1859-1) + 365
  ELSE IF (MOD (jyear-1,100) .NE. 0) THEN
    time_BOY (jyear-1859) = time_BOY (jyear-1859-1) + 366
  ELSE IF (MOD (jyear-1,400) .NE. 0) THEN
    time_BOY (jyear-1859) = time_BOY (jyear-1859-1) + 365
  ELSE
    time_BOY (jyear-1859) = time_BOY (jyear-1859-1) + 366
  END IF
END DO
axy_npp = zero / zero ; axy_plant_mass = zero / zero ; axy_rnf = zero / zero
axy_evap = zero / zero ; axy_theta = zero / zero ; axy_theta_total = zero
!----------------------------------------------------------------------!
! The loop nest of HYBRID9.f90:93-295, calling the reference's routines.
!----------------------------------------------------------------------!
DO iDEC = iDEC_start, iDEC_end
  syr = (iDEC - 1) * 10 + 1901
  IF (iDEC < 12) THEN
    eyr = syr + 9
  ELSE
    eyr = syr + 1
  END IF
  NTIMES = time_BOY (eyr+1-1859) - time_BOY (syr-1859)
  ALLOCATE (tas (lon_c,lat_c,NTIMES), rlds (lon_c,lat_c,NTIMES), rsds (lon_c,lat_c,NTIMES))
  ALLOCATE (huss (lon_c,lat_c,NTIMES), ps (lon_c,lat_c,NTIMES), pr (lon_c,lat_c,NTIMES))
  ALLOCATE (rhs (lon_c,lat_c,NTIMES))
  WRITE (tag,'(A2,I2.2)') 'ec', iDEC
  CALL READ_R (TRIM (dir) // '/tas_d'  // tag // '.f32', tas,  lon_c * lat_c * NTIMES)
  CALL READ_R (TRIM (dir) // '/rlds_d' // tag // '.f32', rlds, lon_c * lat_c * NTIMES)
  CALL READ_R (TRIM (dir) // '/rsds_d' // tag // '.f32', rsds, lon_c * lat_c * NTIMES)
  CALL READ_R (TRIM (dir) // '/huss_d' // tag // '.f32', huss, lon_c * lat_c * NTIMES)
  CALL READ_R (TRIM (dir) // '/ps_d'   // tag // '.f32', ps,   lon_c * lat_c * NTIMES)
  CALL READ_R (TRIM (dir) // '/pr_d'   // tag // '.f32', pr,   lon_c * lat_c * NTIMES)
  CALL READ_R (TRIM (dir) // '/rhs_d'  // tag // '.f32', rhs,  lon_c * lat_c * NTIMES)
  DO y = 1, lat_c
    DO x = 1, lon_c
      IF ((soil_tex (x,y) > 0) .AND. (soil_tex (x,y) /= 13) .AND. &
        SUM (theta_s (:,x,y)) > trunc) THEN
        nlayers = nsoil_layers_max
        IF (.NOT. LEAK) smp (:) = smp_cell (:,x,y)
        DO jyear = syr, eyr
          npp_sum = zero ; plant_mass_sum = zero ; rnf_sum = zero ; evap_sum = zero
          h2osoi_sum_total = zero ; theta_sum (:) = zero
          DO iTIME = time_BOY (jyear-1859), time_BOY (jyear+1-1859) - 1
            iT = iTIME - time_BOY (syr-1859) + 1
            DOY = iTIME - time_BOY (jyear-1859) + 1
            tak = tas (x,y,iT)
            rh = rhs (x,y,iT)
            Rnet = 0.92 * rsds (x,y,iT) + rlds (x,y,iT) - stbo * tas (x,y,iT) ** 4
            PAR = 0.92 * rsds (x,y,iT) * 2.3
            ppt = pr (x,y,iT)
            forc_rain = 1.0E3 * pr (x,y,iT) / rhow
            lamb = ((2503.0 - 2.386 * (tak - tf))) * 1.0E3
            evap_day = zero
            evap_grnd_day = zero
            DO NS = 1, NISURF
              CALL HYDROLOGY
              evap_day = evap_day + (qflx_evap_grnd + qflx_tran_veg_col) * dt
              evap_grnd_day = evap_grnd_day + qflx_evap_grnd * dt
            END DO
            CALL GROW
            DO K = 1, nplants (x,y)
              plant_mass_sum = plant_mass_sum + plant_mass (K,x,y)
            END DO
            npp_sum = npp_sum + npp
            DO I = 1, nlayers
              theta_sum (I) = theta_sum (I) + theta (I)
              h2osoi_sum_total = h2osoi_sum_total + h2osoi_liq (I,x,y)
            END DO
          END DO
          nt = (time_BOY (jyear + 1 - 1859) - 1) - (time_BOY (jyear - 1859)) + 1
          iY = jyear-((iDEC_start-1)*10+1901)+1
          axy_npp (x,y,iY) = npp_sum
          axy_plant_mass (x,y,iY) = plant_mass_sum / FLOAT (nt)
          axy_rnf  (x,y,iY) = rnf_sum  / FLOAT (nt * NISURF)
          axy_evap (x,y,iY) = evap_sum / FLOAT (nt * NISURF)
          DO I = 1, nlayers
            axy_theta (I,x,y,iY) = theta_sum (I) / FLOAT (nt)
          END DO
          axy_theta_total (x,y,iY) = h2osoi_sum_total / FLOAT (nt)
        END DO
        IF (.NOT. LEAK) smp_cell (:,x,y) = smp (:)
      END IF
    END DO
  END DO
  DEALLOCATE (tas, rlds, rsds, huss, ps, pr, rhs)
END DO
!----------------------------------------------------------------------!
CALL WRITE_R (TRIM (dir) // '/out_ref_axy_npp.f32', axy_npp, lon_c * lat_c * NYR)
CALL WRITE_R (TRIM (dir) // '/out_ref_axy_plant_mass.f32', axy_plant_mass, lon_c * lat_c * NYR)
CALL WRITE_R (TRIM (dir) // '/out_ref_axy_rnf.f32', axy_rnf, lon_c * lat_c * NYR)
CALL WRITE_R (TRIM (dir) // '/out_ref_axy_evap.f32', axy_evap, lon_c * lat_c * NYR)
CALL WRITE_R (TRIM (dir) // '/out_ref_axy_theta_total.f32', axy_theta_total, lon_c * lat_c * NYR)
CALL WRITE_R (TRIM (dir) // '/out_ref_axy_theta.f32', axy_theta, n * lon_c * lat_c * NYR)
CALL WRITE_R (TRIM (dir) // '/out_ref_state_h2osoi_liq.f32', h2osoi_liq, n * lon_c * lat_c)
CALL WRITE_R (TRIM (dir) // '/out_ref_state_zwt.f32', zwt, lon_c * lat_c)
CALL WRITE_R (TRIM (dir) // '/out_ref_state_plant_mass.f32', plant_mass, lon_c * lat_c)
WRITE (*,*) 'h9_ref_driver done'
CONTAINS
SUBROUTINE READ_R (path, a, cnt)
  CHARACTER (LEN = *), INTENT (IN) :: path
  INTEGER, INTENT (IN) :: cnt
  REAL, INTENT (OUT) :: a (cnt)
  OPEN (NEWUNIT = u, FILE = path, ACCESS = 'STREAM', FORM = 'UNFORMATTED', STATUS = 'OLD')
  READ (u) a
  CLOSE (u)
END SUBROUTINE READ_R
SUBROUTINE READ_I (path, a, cnt)
  CHARACTER (LEN = *), INTENT (IN) :: path
  INTEGER, INTENT (IN) :: cnt
  INTEGER, INTENT (OUT) :: a (cnt)
  OPEN (NEWUNIT = u, FILE = path, ACCESS = 'STREAM', FORM = 'UNFORMATTED', STATUS = 'OLD')
  READ (u) a
  CLOSE (u)
END SUBROUTINE READ_I
SUBROUTINE WRITE_R (path, a, cnt)
  CHARACTER (LEN = *), INTENT (IN) :: path
  INTEGER, INTENT (IN) :: cnt
  REAL, INTENT (IN) :: a (cnt)
  OPEN (NEWUNIT = u, FILE = path, ACCESS = 'STREAM', FORM = 'UNFORMATTED', STATUS = 'REPLACE')
  WRITE (u) a
  CLOSE (u)
END SUBROUTINE WRITE_R
END PROGRAM H9_REF_DRIVER
