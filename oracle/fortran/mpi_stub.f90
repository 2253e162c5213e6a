!======================================================================!
! Stub for the one thing CONTROL.f90 takes from module MPI
! (CONTROL.f90:10,83: INTEGER :: status (MPI_STATUS_SIZE)), so that the
! reference's SHARED.f90, CONTROL.f90, HYDROLOGY.f90 and GROW.f90 compile
! UNCHANGED without an MPI installation.  Test infrastructure only.
!======================================================================!
MODULE MPI
IMPLICIT NONE
INTEGER, PARAMETER :: MPI_STATUS_SIZE = 6
END MODULE MPI
